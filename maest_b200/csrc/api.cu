// C-ABI entry points of libmaest_b200.so (see include/maest_b200.h for the contract and reference citations).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/maest_b200.h"
#include "attention.cuh"
#include "gemm.cuh"
#include "logmel.cuh"
#include "logmel_tables.h"
#include "rowops.cuh"
#include "tokens.cuh"

using namespace mb;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return fail(-10, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
LogMelTables* g_lm_tables[64] = {nullptr};
bool g_inited[64] = {false};
int g_num_sms[64] = {0};

int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

// 2-D tensor map over a row-major 16-bit matrix [rows, cols] (row stride ld elements); box = [box_rows, 64 cols]
// (128-byte inner extent), SWIZZLE_128B, out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* m, const void* ptr, int dt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  if (!g_encode) return fail(-3, "maest_init() was not called");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16) return fail(-4, "operand not 16-byte aligned (ptr %p ld %llu)", ptr, (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, dt == MAEST_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-5, "cuTensorMapEncodeTiled failed with %d (rows %llu cols %llu ld %llu)", int(r),
                                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}

template <int DT, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  const int num_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + GEMM_BN - 1) / GEMM_BN);
  const int sms = g_num_sms[cur_device()];
  const int grid = num_tiles < sms ? num_tiles : sms;
  gemm_tn_kernel<DT, EPI><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(ta, tb, p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <int DT>
int launch_gemm_dt(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case MAEST_EPI_STORE16: return launch_gemm<DT, EPI_STORE16>(ta, tb, p, st);
    case MAEST_EPI_GELU16: return launch_gemm<DT, EPI_GELU16>(ta, tb, p, st);
    case MAEST_EPI_RESID32: return launch_gemm<DT, EPI_RESID32>(ta, tb, p, st);
    case MAEST_EPI_STORE32: return launch_gemm<DT, EPI_STORE32>(ta, tb, p, st);
  }
  return fail(-1, "unknown epilogue %d", epi);
}

template <typename K>
int set_smem(K kernel, int bytes) {
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

template <int DT>
int init_dt() {
  int r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE16>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_GELU16>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_RESID32>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE32>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_kernel<DT, true>, att_smem_bytes<true>()))) return r;
  if ((r = set_smem(attention_fwd_kernel<DT, false>, att_smem_bytes<false>()))) return r;
  return 0;
}

}  // namespace

extern "C" {

const char* maest_last_error(void) { return g_err; }
int32_t maest_abi_version(void) { return 1; }

int32_t maest_init(int32_t device) {
  if (device < 0 || device >= 64) return fail(-1, "bad device %d", device);
  CUDA_OK(cudaSetDevice(device));
  if (g_inited[device]) return 0;
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(-6, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  g_num_sms[device] = prop.multiProcessorCount;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(-7, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int r;
  if ((r = init_dt<DT_F16>())) return r;
  if ((r = init_dt<DT_BF16>())) return r;
  if ((r = set_smem(logmel_kernel, LM_SMEM_BYTES))) return r;
  static LogMelTables host_tb;
  if (build_logmel_tables(&host_tb)) return fail(-8, "mel filterbank band wider than %d bins", LM_MAX_TAPS);
  CUDA_OK(cudaMalloc(&g_lm_tables[device], sizeof(LogMelTables)));
  CUDA_OK(cudaMemcpy(g_lm_tables[device], &host_tb, sizeof(LogMelTables), cudaMemcpyHostToDevice));
  g_inited[device] = true;
  return 0;
}

int32_t maest_logmel_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, float* mel, void* stream) {
  const int dev = cur_device();
  if (!g_inited[dev]) return fail(-3, "maest_init() was not called for device %d", dev);
  if (B <= 0) return 0;
  if (S <= LM_HOP) return fail(-1, "waveform of %d samples is too short for reflect padding (need > 256)", S);
  LogMelParams p;
  p.wav = wav; p.wav_stride = wav_stride; p.B = B; p.S = S; p.T = 1 + S / LM_HOP; p.mel = mel;
  p.tables = g_lm_tables[dev];
  dim3 grid((p.T + LM_FRAMES - 1) / LM_FRAMES, B);
  logmel_kernel<<<grid, LM_THREADS, LM_SMEM_BYTES, (cudaStream_t)stream>>>(p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_linear_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias, int32_t M,
                         int32_t N, int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out,
                         const float* resid, const float* addend, int32_t rows_per_group, int32_t group_stride,
                         int32_t row_offset, void* stream) {
  if (M <= 0) return 0;
  if (N % 32 || K % 8) return fail(-1, "linear: N %% 32 and K %% 8 must be 0 (N %d K %d)", N, K);
  if (op_dtype != MAEST_F16 && op_dtype != MAEST_BF16) return fail(-1, "linear: op_dtype must be f16/bf16");
  CUtensorMap ta, tb;
  int r;
  if ((r = make_tmap(&ta, a, op_dtype, M, K, lda, GEMM_BM))) return r;
  if ((r = make_tmap(&tb, w, op_dtype, N, K, ldw, GEMM_BN))) return r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.resid = resid; p.addend = addend; p.ld_out = int(ld_out);
  if (rows_per_group <= 0) { p.rows_per_group = 0x7fffffff; p.group_stride = 0; p.row_offset = 0; }
  else { p.rows_per_group = rows_per_group; p.group_stride = group_stride; p.row_offset = row_offset; }
  if (epilogue == MAEST_EPI_RESID32 && !resid) return fail(-1, "linear: RESID32 needs resid");
  if (epilogue == MAEST_EPI_RESID32 && rows_per_group > 0) return fail(-1, "linear: RESID32 does not support row remapping");
  cudaStream_t st = (cudaStream_t)stream;
  return op_dtype == MAEST_BF16 ? launch_gemm_dt<DT_BF16>(epilogue, ta, tb, p, st)
                                : launch_gemm_dt<DT_F16>(epilogue, ta, tb, p, st);
}

int32_t maest_layernorm_fwd(const float* x, const float* w, const float* b, void* y16, int32_t op_dtype, int32_t rows,
                            float eps, float* mean, float* rstd, void* stream) {
  if (rows <= 0) return 0;
  const int blocks = (rows + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == MAEST_BF16) layernorm_to16_kernel<DT_BF16><<<blocks, 256, 0, st>>>(x, w, b, y16, rows, eps, mean, rstd);
  else if (op_dtype == MAEST_F16) layernorm_to16_kernel<DT_F16><<<blocks, 256, 0, st>>>(x, w, b, y16, rows, eps, mean, rstd);
  else return fail(-1, "layernorm: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_attention_fwd(const void* qkv, void* out, int32_t B, int32_t N, int32_t H, int32_t op_dtype,
                            int32_t variant, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  CUtensorMap tq;
  int r;
  if ((r = make_tmap(&tq, qkv, op_dtype, uint64_t(B) * N, uint64_t(3) * H * ATT_D, uint64_t(3) * H * ATT_D, 128))) return r;
  AttnParams p;
  p.B = B; p.N = N; p.H = H; p.ld_qkv = 3 * H * ATT_D; p.ld_out = H * ATT_D; p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  dim3 grid((N + ATT_BQ - 1) / ATT_BQ, H, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == MAEST_BF16) {
    if (variant == 0) attention_fwd_kernel<DT_BF16, true><<<grid, ATT_THREADS, att_smem_bytes<true>(), st>>>(tq, p);
    else attention_fwd_kernel<DT_BF16, false><<<grid, ATT_THREADS, att_smem_bytes<false>(), st>>>(tq, p);
  } else if (op_dtype == MAEST_F16) {
    if (variant == 0) attention_fwd_kernel<DT_F16, true><<<grid, ATT_THREADS, att_smem_bytes<true>(), st>>>(tq, p);
    else attention_fwd_kernel<DT_F16, false><<<grid, ATT_THREADS, att_smem_bytes<false>(), st>>>(tq, p);
  } else return fail(-1, "attention: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

size_t maest_patch_workspace_bytes(int32_t B, int32_t P) {
  size_t a16 = (size_t(B) * P * 256 * 2 + 255) & ~size_t(255);
  size_t pos = (size_t(P) * 768 * 4 + 255) & ~size_t(255);
  return a16 + pos;
}

int32_t maest_patch_tokens_fwd(const void* mel, int32_t mel_dtype, int32_t B, int32_t T, const void* w_pe,
                               int32_t op_dtype, const float* conv_bias, const float* freq_pe, int32_t Fp,
                               const float* time_pe, int32_t Wt, const float* cls_token, const float* dist_token,
                               const float* new_pos_embed, const int32_t* keep_ft, int32_t P, int32_t t_offset,
                               float* tokens, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return 0;
  if (T < 16) return fail(-1, "patch_tokens: T=%d shorter than one patch", T);
  const int Tp = (T - 16) / 10 + 1;
  if (Tp + t_offset > Wt) return fail(-2, "the patches shape (time %d + offset %d) is larger than the expected time encodings %d, please reduce the input duration", Tp, t_offset, Wt);
  if (!keep_ft && P != Fp * Tp) return fail(-1, "patch_tokens: P=%d but grid is %d x %d", P, Fp, Tp);
  if (mel_dtype != MAEST_F32 && mel_dtype != MAEST_F16) return fail(-1, "patch_tokens: mel must be f32 or f16");
  if (workspace_bytes < maest_patch_workspace_bytes(B, P)) return fail(-1, "patch_tokens: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* a16 = ws;
  float* pos = reinterpret_cast<float*>(ws + ((size_t(B) * P * 256 * 2 + 255) & ~size_t(255)));

  PosTableParams pp;
  pp.conv_bias = conv_bias; pp.freq_pe = freq_pe; pp.time_pe = time_pe; pp.Fp = Fp; pp.Wt = Wt; pp.Tp = Tp; pp.P = P;
  pp.t_off = t_offset; pp.keep_ft = keep_ft; pp.pos = pos;
  if (P > 0) pos_table_kernel<<<P, 256, 0, st>>>(pp);
  CUDA_OK(cudaGetLastError());

  PatchGatherParams gp;
  gp.mel = mel; gp.mel_is_half = mel_dtype == MAEST_F16; gp.B = B; gp.T = T; gp.P = P; gp.Tp = Tp; gp.keep_ft = keep_ft;
  gp.a16 = a16; gp.cls_token = cls_token; gp.dist_token = dist_token; gp.new_pos_embed = new_pos_embed; gp.tokens = tokens;
  const long warps = long(B) * (P > 2 ? P : 2);
  const int blocks = int((warps + 7) / 8);
  if (op_dtype == MAEST_BF16) patch_gather_kernel<DT_BF16><<<blocks, 256, 0, st>>>(gp);
  else patch_gather_kernel<DT_F16><<<blocks, 256, 0, st>>>(gp);
  CUDA_OK(cudaGetLastError());
  if (P == 0) return 0;
  return maest_linear_fwd(a16, 256, w_pe, 256, nullptr, B * P, 768, 256, op_dtype, MAEST_EPI_STORE32, tokens, 768, nullptr,
                          pos, P, 2 + P, 2, stream);
}

size_t maest_encoder_workspace_bytes(int64_t rows) {
  // h16 [rows,768] | qkv16 [rows,2304] | o16 [rows,768] | u16 [rows,3072]  (+ 128 rows of slack per buffer)
  return size_t(rows + 128) * (768 + 2304 + 768 + 3072) * 2 + 1024;
}

int32_t maest_encoder_fwd(float* x, int32_t B, int32_t N, const MaestBlockWeights* blocks, int32_t n_blocks,
                          int32_t last_attn_only, int32_t op_dtype, int32_t attn_variant, void* workspace,
                          size_t workspace_bytes, void* stream) {
  const int64_t M = int64_t(B) * N;
  if (M <= 0 || n_blocks <= 0) return 0;
  if (workspace_bytes < maest_encoder_workspace_bytes(M)) return fail(-1, "encoder: workspace too small (%zu < %zu)", workspace_bytes, maest_encoder_workspace_bytes(M));
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const size_t R = size_t(M + 128);
  uint8_t* h16 = ws;
  uint8_t* qkv16 = h16 + R * 768 * 2;
  uint8_t* o16 = qkv16 + R * 2304 * 2;
  uint8_t* u16 = o16 + R * 768 * 2;
  int r;
  for (int i = 0; i < n_blocks; ++i) {
    const MaestBlockWeights& w = blocks[i];
    const bool attn_only = last_attn_only && i == n_blocks - 1;
    if ((r = maest_layernorm_fwd(x, w.ln1_w, w.ln1_b, h16, op_dtype, int(M), 1e-6f, nullptr, nullptr, stream))) return r;
    if ((r = maest_linear_fwd(h16, 768, w.qkv_w, 768, w.qkv_b, int(M), 2304, 768, op_dtype, MAEST_EPI_STORE16, qkv16, 2304,
                              nullptr, nullptr, 0, 0, 0, stream))) return r;
    if ((r = maest_attention_fwd(qkv16, o16, B, N, 12, op_dtype, attn_variant, stream))) return r;
    if (attn_only) {
      return maest_linear_fwd(o16, 768, w.proj_w, 768, w.proj_b, int(M), 768, 768, op_dtype, MAEST_EPI_STORE32, x, 768, nullptr,
                              nullptr, 0, 0, 0, stream);
    }
    if ((r = maest_linear_fwd(o16, 768, w.proj_w, 768, w.proj_b, int(M), 768, 768, op_dtype, MAEST_EPI_RESID32, x, 768, x,
                              nullptr, 0, 0, 0, stream))) return r;
    if ((r = maest_layernorm_fwd(x, w.ln2_w, w.ln2_b, h16, op_dtype, int(M), 1e-6f, nullptr, nullptr, stream))) return r;
    if ((r = maest_linear_fwd(h16, 768, w.fc1_w, 768, w.fc1_b, int(M), 3072, 768, op_dtype, MAEST_EPI_GELU16, u16, 3072, nullptr,
                              nullptr, 0, 0, 0, stream))) return r;
    if ((r = maest_linear_fwd(u16, 3072, w.fc2_w, 3072, w.fc2_b, int(M), 768, 3072, op_dtype, MAEST_EPI_RESID32, x, 768, x,
                              nullptr, 0, 0, 0, stream))) return r;
  }
  return 0;
}

int32_t maest_pool_head_fwd(const float* x, int32_t B, int32_t N, const float* norm_w, const float* norm_b,
                            const float* head_ln_w, const float* head_ln_b, const float* head_w, const float* head_b,
                            const float* head_dist_w, const float* head_dist_b, int32_t C, int32_t mode,
                            float* logits, float* logits_dist, float* feats, float* ln_cls, float* ln_dist,
                            void* stream) {
  if (B <= 0) return 0;
  if (mode == MAEST_HEAD_SEPARATED && (!head_dist_w || !head_dist_b || !logits_dist)) return fail(-1, "pool_head: separated mode needs head_dist and logits_dist");
  HeadParams p;
  p.x = x; p.N = N; p.norm_w = norm_w; p.norm_b = norm_b; p.hln_w = head_ln_w; p.hln_b = head_ln_b; p.head_w = head_w;
  p.head_b = head_b; p.hdist_w = head_dist_w; p.hdist_b = head_dist_b; p.C = C; p.separated = mode == MAEST_HEAD_SEPARATED;
  p.logits = logits; p.logits_dist = logits_dist; p.feats = feats; p.ln_cls = ln_cls; p.ln_dist = ln_dist;
  pool_head_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_block_embedding_fwd(const float* x, int32_t B, int32_t N, float* emb, void* stream) {
  if (B <= 0) return 0;
  if (N < 3) return fail(-1, "block_embedding: need at least one patch token");
  block_embedding_kernel<<<dim3(B, 768 / 64), 256, 0, (cudaStream_t)stream>>>(x, N, emb);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_cast_to16(const float* src, void* dst, int64_t n, int32_t op_dtype, void* stream) {
  if (n <= 0) return 0;
  const long blocks = (n + 1023) / 1024;
  if (op_dtype == MAEST_BF16) cast_to16_kernel<DT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  else if (op_dtype == MAEST_F16) cast_to16_kernel<DT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  else return fail(-1, "cast: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
