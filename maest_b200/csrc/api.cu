// C-ABI entry points of libmaest_b200.so (see include/maest_b200.h for the contract and reference citations).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/maest_b200.h"
#include "attention.cuh"
#include "attention_bwd.cuh"
#include "attention_chain.cuh"
#include "ingest.cuh"
#include "metrics.cuh"
#include "optim.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "logmel.cuh"
#include "logmel_tables.h"
#include "rowops.cuh"
#include "tokens.cuh"
#include "train.cuh"

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
// A tensor map is a pure function of (pointer, dtype, shape, stride, box): the activations live in caller-owned buffers that
// are reused step after step and the weights never move, so the ~200 encodes per forward step collapse into lookups of a small
// per-thread set-associative cache (cuTensorMapEncodeTiled costs a microsecond or two on the host -- invisible at batch 64,
// the dominant host cost of a one-clip predict_labels call).
struct TmapKey {
  const void* ptr; uint64_t rows, cols, ld; uint32_t box_rows; int32_t dt;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && dt == o.dt;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
// 4-way set-associative, round-robin replacement.  (A direct-mapped table of 1024 slots thrashed whenever two of the ~60 live
// descriptors of a forward hashed to the same slot -- about 1.8 colliding pairs per process on average, depending on where the
// allocator happened to put the buffers: a run-to-run lottery for tests/test_gpu_parity.py::test_tensor_map_cache_serves_repeat_calls.)
constexpr int TMAP_CACHE_SETS = 512, TMAP_CACHE_WAYS = 4;
thread_local TmapSlot g_tmap_cache[TMAP_CACHE_SETS * TMAP_CACHE_WAYS];
thread_local uint8_t g_tmap_next_way[TMAP_CACHE_SETS];
uint64_t g_tmap_hits = 0, g_tmap_misses = 0;

int make_tmap_uncached(CUtensorMap* m, const void* ptr, int dt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int make_tmap_f32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

int make_tmap(CUtensorMap* m, const void* ptr, int dt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  const TmapKey key{ptr, rows, cols, ld, box_rows, dt};
  uint64_t h = reinterpret_cast<uintptr_t>(ptr) >> 4;
  h = (h ^ (h >> 17) ^ (rows * 0x9E3779B97F4A7C15ull) ^ (cols << 7) ^ (ld << 13) ^ (uint64_t(box_rows) << 29) ^ uint64_t(dt)) * 0xD6E8FEB86659FD93ull;
  const int set = int((h >> 40) % TMAP_CACHE_SETS);
  TmapSlot* ways = &g_tmap_cache[set * TMAP_CACHE_WAYS];
  static const bool enabled = [] { const char* e = getenv("MAEST_TMAP_CACHE"); return !(e && e[0] == '0'); }();   // A/B switch
  if (enabled) {
    for (int w = 0; w < TMAP_CACHE_WAYS; ++w) {
      if (ways[w].valid && ways[w].key == key) {
        *m = ways[w].map;
        ++g_tmap_hits;
        return 0;
      }
    }
  }
  const int r = dt == MAEST_F32 ? make_tmap_f32(m, ptr, rows, cols, ld, box_rows) : make_tmap_uncached(m, ptr, dt, rows, cols, ld, box_rows);
  if (r == 0) {
    int w = 0;
    while (w < TMAP_CACHE_WAYS && ways[w].valid) ++w;                      // a free way first, else round-robin
    if (w == TMAP_CACHE_WAYS) { w = g_tmap_next_way[set]; g_tmap_next_way[set] = uint8_t((w + 1) % TMAP_CACHE_WAYS); }
    ways[w].key = key; ways[w].map = *m; ways[w].valid = true;
    ++g_tmap_misses;
  }
  return r;
}

int make_tmap_uncached(CUtensorMap* m, const void* ptr, int dt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
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

// fp32 [rows, cols] matrix, box = [box_rows, 32 floats] (128-byte inner extent), SWIZZLE_128B: target of the TMA reduce-add
int make_tmap_f32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  if (!g_encode) return fail(-3, "maest_init() was not called");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-5, "cuTensorMapEncodeTiled(f32) failed with %d", int(r));
  return 0;
}

int g_gemm_pair_mode = 2;   // 0: 1-CTA tiles, 1: CTA-pair (cta_group::2) tiles, 2: per-shape choice (default)

// Measured on B200 (M = 107 840): the pair kernel wins where the mainloop dominates (qkv 0.329 -> 0.289 ms, fc2 0.434 -> 0.400 ms)
// and, with the packed / TMA-store epilogue, for the inference fc1 + GELU (0.476 -> 0.450 ms); it loses for proj (0.188 -> 0.197 ms)
// and, once both of its outputs left through TMA stores, also for the training forward fc1 (GELU16_SAVE).
inline bool use_pair_kernel(int epi, int K, bool has_aux = false) {
  if (g_gemm_pair_mode != 2) return g_gemm_pair_mode == 1;
  return epi == MAEST_EPI_STORE16 || epi == MAEST_EPI_STORE16_LN || epi == MAEST_EPI_GELU16 || epi == MAEST_EPI_GELU16_LN ||
         ((epi == MAEST_EPI_RESID32 || epi == MAEST_EPI_RESID32_LN) && K >= 2048);
}

// Tensor map of the 16-bit output for the TMA-store epilogues (STORE16 / GELU16 and their LN-folded forms); set by the entry
// points right before the dispatch, ignored by every other epilogue.
static thread_local CUtensorMap t_tmap_c, t_tmap_c2;   // (second map: the saved pre-activation of GELU16_SAVE)

template <int DT, int EPI>
int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  const int num_tiles = ((p.M + 255) / 256) * ((p.N + GEMM_BN - 1) / GEMM_BN);
  const int sms = g_num_sms[cur_device()];
  int pairs = sms / 2;
  if (pairs > num_tiles) pairs = num_tiles;
  gemm2_tn_kernel<DT, EPI><<<2 * pairs, GEMM_THREADS, GEMM2_SMEM_BYTES, st>>>(ta, tb, t_tmap_c, t_tmap_c2, p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <int DT, int EPI, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  const int splits = p.k_splits > 1 ? p.k_splits : 1;
  const int num_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + GEMM_BN - 1) / GEMM_BN) * splits;
  const int sms = g_num_sms[cur_device()];
  const int grid = num_tiles < sms ? num_tiles : sms;
  gemm_tn_kernel<DT, EPI, A_MN, B_MN><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(ta, tb, t_tmap_c, t_tmap_c2, p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// the instantiated (epilogue, operand-major) combinations: forward (K,K), dgrad (K,MN), wgrad (MN,MN)
template <int DT>
int launch_gemm_dt(int epi, bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn && p.k_splits <= 1 && use_pair_kernel(epi, p.K, p.aux16 != nullptr)) {
    switch (epi) {
      case MAEST_EPI_STORE16: return launch_gemm2<DT, EPI_STORE16>(ta, tb, p, st);
      case MAEST_EPI_GELU16:
        if (p.aux16 != nullptr) return launch_gemm2<DT, EPI_GELU16_SAVE>(ta, tb, p, st);
        return launch_gemm2<DT, EPI_GELU16>(ta, tb, p, st);
      case MAEST_EPI_RESID32: return launch_gemm2<DT, EPI_RESID32>(ta, tb, p, st);
      case MAEST_EPI_STORE32: return launch_gemm2<DT, EPI_STORE32>(ta, tb, p, st);
      case MAEST_EPI_STORE16_LN: return launch_gemm2<DT, EPI_STORE16_LN>(ta, tb, p, st);
      case MAEST_EPI_GELU16_LN: return launch_gemm2<DT, EPI_GELU16_LN>(ta, tb, p, st);
      case MAEST_EPI_RESID32_LN: return launch_gemm2<DT, EPI_RESID32_LN>(ta, tb, p, st);
    }
  }
  if (!a_mn && !b_mn) {
    switch (epi) {
      case MAEST_EPI_STORE16: return launch_gemm<DT, EPI_STORE16, false, false>(ta, tb, p, st);
      case MAEST_EPI_GELU16:
        if (p.aux16 != nullptr) return launch_gemm<DT, EPI_GELU16_SAVE, false, false>(ta, tb, p, st);
        return launch_gemm<DT, EPI_GELU16, false, false>(ta, tb, p, st);
      case MAEST_EPI_RESID32: return launch_gemm<DT, EPI_RESID32, false, false>(ta, tb, p, st);
      case MAEST_EPI_STORE32: return launch_gemm<DT, EPI_STORE32, false, false>(ta, tb, p, st);
      case MAEST_EPI_STORE16_LN: return launch_gemm<DT, EPI_STORE16_LN, false, false>(ta, tb, p, st);
      case MAEST_EPI_GELU16_LN: return launch_gemm<DT, EPI_GELU16_LN, false, false>(ta, tb, p, st);
      case MAEST_EPI_RESID32_LN: return launch_gemm<DT, EPI_RESID32_LN, false, false>(ta, tb, p, st);
    }
  } else if (!a_mn && b_mn) {
    switch (epi) {
      case MAEST_EPI_STORE16: return launch_gemm<DT, EPI_STORE16, false, true>(ta, tb, p, st);
      case MAEST_EPI_STORE32: return launch_gemm<DT, EPI_STORE32, false, true>(ta, tb, p, st);
      case MAEST_EPI_GELUBWD16: return launch_gemm<DT, EPI_GELUBWD16, false, true>(ta, tb, p, st);
    }
  } else if (a_mn && b_mn) {
    if (epi == MAEST_EPI_ATOMIC32) return launch_gemm<DT, EPI_ATOMIC32, true, true>(ta, tb, p, st);
  }
  return fail(-1, "gemm: epilogue %d is not built for operand majors (a_mn %d, b_mn %d)", epi, int(a_mn), int(b_mn));
}

template <typename K>
int set_smem(K kernel, int bytes) {
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

template <int DT>
int init_dt() {
  int r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE16, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_GELU16, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_GELU16_SAVE, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_RESID32, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE32, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE16, false, true>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE32, false, true>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_GELUBWD16, false, true>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_ATOMIC32, true, true>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_STORE16>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_GELU16>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_GELU16_SAVE>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_RESID32>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_STORE32>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_STORE16_LN>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_GELU16_LN>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm2_tn_kernel<DT, EPI_RESID32_LN>, GEMM2_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_STORE16_LN, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_GELU16_LN, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(gemm_tn_kernel<DT, EPI_RESID32_LN, false, false>, GEMM_SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_kernel<DT, true>, att_smem_bytes<true>()))) return r;
  if ((r = set_smem(attention_fwd_spec_kernel<DT, 128>, 120 * 1024))) return r;
  if ((r = set_smem(attention_fwd_kernel<DT, false>, att_smem_bytes<false>()))) return r;
  if ((r = set_smem(attention_bwd_kernel<DT>, ATTB_SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg3>, AtcCfg3::SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg4>, AtcCfg4::SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg3x2>, AtcCfg3x2::SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg3L>, AtcCfg3L::SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg3E>, AtcCfg3E::SMEM_BYTES))) return r;
  if ((r = set_smem(attention_fwd_chain_kernel<DT, AtcCfg3EH>, AtcCfg3EH::SMEM_BYTES))) return r;
  return 0;
}

}  // namespace

extern "C" {

const char* maest_last_error(void) { return g_err; }
int32_t maest_abi_version(void) { return 8; }

int32_t maest_tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  if (hits) *hits = g_tmap_hits;
  if (misses) *misses = g_tmap_misses;
  return 0;
}

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

static int32_t logmel_launch(const float* wav, int32_t B, int32_t S, int64_t wav_stride, float* mel, void* raw_tm16, void* stream,
                             int32_t essentia_framing = 0) {
  const int dev = cur_device();
  if (!g_inited[dev]) return fail(-3, "maest_init() was not called for device %d", dev);
  if (B <= 0) return 0;
  if (S <= LM_HOP) return fail(-1, "waveform of %d samples is too short for reflect padding (need > 256)", S);
  LogMelParams p;
  p.wav = wav; p.wav_stride = wav_stride; p.B = B; p.S = S; p.mel = mel;
  p.T = essentia_framing ? (S + LM_HOP - 1) / LM_HOP : 1 + S / LM_HOP;
  p.essentia_framing = essentia_framing ? 1 : 0;
  p.raw_tm16 = reinterpret_cast<__half*>(raw_tm16);
  p.tables = g_lm_tables[dev];
  dim3 grid((p.T + LM_FRAMES - 1) / LM_FRAMES, B);
  logmel_kernel<<<grid, LM_THREADS, LM_SMEM_BYTES, (cudaStream_t)stream>>>(p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_logmel_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, float* mel, void* stream) {
  if (!mel) return fail(-1, "logmel: mel is NULL");
  return logmel_launch(wav, B, S, wav_stride, mel, nullptr, stream);
}

int32_t maest_logmel_raw16_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, void* raw_tm16, int32_t framing, void* stream) {
  if (!raw_tm16) return fail(-1, "logmel_raw16: output is NULL");
  if (framing != 0 && framing != 1) return fail(-1, "logmel_raw16: framing must be 0 (torchaudio) or 1 (essentia), got %d", framing);
  return logmel_launch(wav, B, S, wav_stride, nullptr, raw_tm16, stream, framing);
}

int32_t maest_adamw_step(const void* tensor_table, const void* chunk_table, int32_t n_chunks, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int32_t step, float grad_scale, float swa_inv, void* stream) {
  if (n_chunks <= 0) return 0;
  if (step < 1) return fail(-1, "adamw: step counts from 1");
  AdamWParams a;
  a.tensors = reinterpret_cast<const OptTensor*>(tensor_table);
  a.chunks = reinterpret_cast<const OptChunk*>(chunk_table);
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_c1 = float(1.0 - pow(double(beta1), double(step)));
  a.bias_c2_sqrt = float(sqrt(1.0 - pow(double(beta2), double(step))));
  a.grad_scale = grad_scale; a.swa_inv = swa_inv;
  adamw_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(a);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_swa_fold(const void* tensor_table, const void* chunk_table, int32_t n_chunks, float swa_inv, void* stream) {
  if (n_chunks <= 0) return 0;
  swa_fold_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const OptTensor*>(tensor_table),
                                                                     reinterpret_cast<const OptChunk*>(chunk_table), swa_inv);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_ap_roc_fwd(const float* score_sorted, const float* label_sorted, int32_t n, int32_t C, double* ap, double* auc,
                         int32_t* n_pos, void* stream) {
  if (n <= 0 || C <= 0) return 0;
  ap_roc_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(score_sorted, label_sorted, n, C, ap, auc, n_pos);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_mel_ingest_fwd(const void* raw_tm16, const int32_t* frames_read, const int32_t* roll_shift, int32_t B, int32_t T,
                             int32_t do_norm, float norm_mean, float norm_std, void* out, void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(raw_tm16) & 3) != 0) return fail(-4, "mel_ingest: raw windows must be 4-byte aligned");
  IngestParams p;
  p.raw = reinterpret_cast<const __half*>(raw_tm16); p.frames_read = frames_read; p.roll_shift = roll_shift; p.B = B; p.T = T;
  p.do_norm = do_norm;
  // the reference's arithmetic: float16 array (op) Python float -> numpy rounds the scalar to float16 first
  p.norm_mean = __float2half_rn(norm_mean);
  p.norm_2std = __float2half_rn(norm_std * 2.0f);
  p.out = reinterpret_cast<__half*>(out);
  dim3 grid((T + ING_FRAMES - 1) / ING_FRAMES, B);
  mel_ingest_kernel<<<grid, ING_THREADS, 0, (cudaStream_t)stream>>>(p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_gemm(const void* a, int64_t lda, int32_t a_mn, const void* b, int64_t ldb, int32_t b_mn, const float* bias,
                   int32_t M, int32_t N, int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out,
                   const float* resid, const float* addend, int32_t rows_per_group, int32_t group_stride,
                   int32_t row_offset, void* aux16, int32_t k_splits, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (N % 32) return fail(-1, "gemm: N %% 32 must be 0 (M %d N %d K %d)", M, N, K);   // operand alignment is checked per tensor map
  if (op_dtype != MAEST_F16 && op_dtype != MAEST_BF16) return fail(-1, "gemm: op_dtype must be f16/bf16");
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  if (k_splits > num_kb) k_splits = num_kb;
  if (k_splits > 1 && epilogue != MAEST_EPI_ATOMIC32) return fail(-1, "gemm: split-K needs the ATOMIC32 epilogue");
  CUtensorMap ta, tb;
  int r;
  // K-major operand: matrix [rows = M|N, cols = K]; MN-major operand: matrix [rows = K, cols = M|N]
  if ((r = a_mn ? make_tmap(&ta, a, op_dtype, K, M, lda, 64) : make_tmap(&ta, a, op_dtype, M, K, lda, GEMM_BM))) return r;
  const bool pair_kernel = !a_mn && !b_mn && k_splits <= 1 && use_pair_kernel(epilogue, K, aux16 != nullptr);   // each CTA of a pair loads half of the W tile
  if ((r = b_mn ? make_tmap(&tb, b, op_dtype, K, N, ldb, 64) : make_tmap(&tb, b, op_dtype, N, K, ldb, pair_kernel ? 128 : GEMM_BN))) return r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.resid = resid; p.addend = addend; p.ld_out = int(ld_out);
  p.aux16 = aux16; p.k_splits = k_splits; p.reduce_out = 0;
  p.ln_stats = nullptr; p.ln_vec = nullptr; p.out16b = nullptr; p.ln_rows = 0;
  if (rows_per_group <= 0) { p.rows_per_group = 0x7fffffff; p.group_stride = 0; p.row_offset = 0; }
  else { p.rows_per_group = rows_per_group; p.group_stride = group_stride; p.row_offset = row_offset; }
  if (epilogue == MAEST_EPI_RESID32 && !resid) return fail(-1, "gemm: RESID32 needs resid");
  if (epilogue == MAEST_EPI_RESID32 && rows_per_group > 0) return fail(-1, "gemm: RESID32 does not support row remapping");
  if ((epilogue == MAEST_EPI_STORE16 || epilogue == MAEST_EPI_GELU16) && rows_per_group > 0) return fail(-1, "gemm: the 16-bit epilogues do not support row remapping");
  if (epilogue == MAEST_EPI_GELUBWD16 && !aux16) return fail(-1, "gemm: GELUBWD16 needs the saved pre-activation (aux16)");
  if (epilogue == MAEST_EPI_GELUBWD16 && bias) return fail(-1, "gemm: GELUBWD16 (an input gradient) takes no bias");
  if ((epilogue == MAEST_EPI_STORE16 || epilogue == MAEST_EPI_GELU16 || epilogue == MAEST_EPI_GELUBWD16) &&
      (r = make_tmap(&t_tmap_c, out, op_dtype, M, N, ld_out, 32))) return r;     // TMA-store epilogue: [32 rows x 64 columns] boxes
  if (epilogue == MAEST_EPI_GELU16 && aux16 && (r = make_tmap(&t_tmap_c2, aux16, op_dtype, M, N, ld_out, 32))) return r;
  // in-place residual update (x += A W^T + b, the inference encoder's proj and fc2): the add is done by a TMA reduce into out
  static const bool reduce_enabled = [] { const char* e = getenv("MAEST_RESID_REDUCE"); return !(e && e[0] == '0'); }();   // A/B switch
  if (epilogue == MAEST_EPI_RESID32 && reduce_enabled && resid == reinterpret_cast<const float*>(out) && (ld_out % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    if ((r = make_tmap(&t_tmap_c, out, MAEST_F32, M, N, ld_out, 32))) return r;
    p.reduce_out = 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return op_dtype == MAEST_BF16 ? launch_gemm_dt<DT_BF16>(epilogue, a_mn != 0, b_mn != 0, ta, tb, p, st)
                                : launch_gemm_dt<DT_F16>(epilogue, a_mn != 0, b_mn != 0, ta, tb, p, st);
}

int32_t maest_linear_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias, int32_t M,
                         int32_t N, int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out,
                         const float* resid, const float* addend, int32_t rows_per_group, int32_t group_stride,
                         int32_t row_offset, void* stream) {
  if (epilogue < MAEST_EPI_STORE16 || epilogue > MAEST_EPI_STORE32) return fail(-1, "linear: unknown epilogue %d", epilogue);
  return maest_gemm(a, lda, 0, w, ldw, 0, bias, M, N, K, op_dtype, epilogue, out, ld_out, resid, addend, rows_per_group,
                    group_stride, row_offset, nullptr, 1, stream);
}

int32_t maest_ln_fold(const void* w16, const float* gamma, const float* beta, const float* bias, int32_t N, int32_t K,
                      int32_t op_dtype, float* wg, float* bf, void* stream) {
  if (N <= 0) return 0;
  const int blocks = (N + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == MAEST_BF16) ln_fold_kernel<DT_BF16><<<blocks, 256, 0, st>>>(w16, gamma, beta, bias, N, K, wg, bf);
  else if (op_dtype == MAEST_F16) ln_fold_kernel<DT_F16><<<blocks, 256, 0, st>>>(w16, gamma, beta, bias, N, K, wg, bf);
  else return fail(-1, "ln_fold: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_ln_finalize(const float* partials, int32_t rows, int32_t n_features, float eps, float* stats, void* stream) {
  if (rows <= 0) return 0;
  if (n_features % GEMM_LN_PART) return fail(-1, "ln_finalize: n_features %% %d must be 0", GEMM_LN_PART);
  ln_finalize_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(partials), rows, rows, n_features / GEMM_LN_PART, eps,
                                                                             reinterpret_cast<float2*>(stats));
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_linear_ln_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias, int32_t M, int32_t N,
                            int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out, const float* resid,
                            float* ln_stats, const float* ln_vec, void* out16b, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (epilogue != MAEST_EPI_STORE16_LN && epilogue != MAEST_EPI_GELU16_LN && epilogue != MAEST_EPI_RESID32_LN)
    return fail(-1, "linear_ln: epilogue %d is not a LayerNorm-folding epilogue", epilogue);
  if (N % 32) return fail(-1, "linear_ln: N %% 32 must be 0");
  if (op_dtype != MAEST_F16 && op_dtype != MAEST_BF16) return fail(-1, "linear_ln: op_dtype must be f16/bf16");
  if (!ln_stats || !ln_vec) return fail(-1, "linear_ln: ln_stats / ln_vec are required");
  if (epilogue == MAEST_EPI_RESID32_LN && (!resid || !out16b)) return fail(-1, "linear_ln: the producer epilogue needs resid and out16b");
  CUtensorMap ta, tb;
  int r;
  if ((r = make_tmap(&ta, a, op_dtype, M, K, lda, GEMM_BM))) return r;
  const bool pair_kernel = use_pair_kernel(epilogue, K);
  if ((r = make_tmap(&tb, w, op_dtype, N, K, ldw, pair_kernel ? 128 : GEMM_BN))) return r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.resid = resid; p.addend = nullptr; p.ld_out = int(ld_out);
  p.aux16 = nullptr; p.k_splits = 1; p.reduce_out = 0; p.rows_per_group = 0x7fffffff; p.group_stride = 0; p.row_offset = 0;
  p.ln_stats = ln_stats; p.ln_vec = ln_vec; p.out16b = out16b; p.ln_rows = M;
  if ((epilogue == MAEST_EPI_STORE16_LN || epilogue == MAEST_EPI_GELU16_LN) && (r = make_tmap(&t_tmap_c, out, op_dtype, M, N, ld_out, 32))) return r;
  cudaStream_t st = (cudaStream_t)stream;
  return op_dtype == MAEST_BF16 ? launch_gemm_dt<DT_BF16>(epilogue, false, false, ta, tb, p, st)
                                : launch_gemm_dt<DT_F16>(epilogue, false, false, ta, tb, p, st);
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

int32_t maest_set_gemm_mode(int32_t pair_mode) {
  g_gemm_pair_mode = pair_mode < 0 || pair_mode > 2 ? 2 : pair_mode;
  return 0;
}

int32_t maest_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t N, int32_t H, int32_t op_dtype,
                            int32_t variant, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(out) & 31) return fail(-4, "attention_fwd: out must be 32-byte aligned (256-bit stores)");
  CUtensorMap tq;
  int r;
  if ((r = make_tmap(&tq, qkv, op_dtype, uint64_t(B) * N, uint64_t(3) * H * ATT_D, uint64_t(3) * H * ATT_D, 128))) return r;
  AttnParams p;
  p.B = B; p.N = N; p.H = H; p.ld_qkv = 3 * H * ATT_D; p.ld_out = H * ATT_D; p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.lse = lse;
  dim3 grid((N + ATT_BQ - 1) / ATT_BQ, H, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype != MAEST_BF16 && op_dtype != MAEST_F16) return fail(-1, "attention: op_dtype must be f16/bf16");
  const bool bf = op_dtype == MAEST_BF16;
  switch (variant) {
    case 0:    // default: speculative running max, peeled last KV tile
      if (bf) attention_fwd_spec_kernel<DT_BF16, 128><<<grid, ATT_THREADS, AttSpecCfg<128>::kSmemLaunch, st>>>(tq, tq, p);
      else attention_fwd_spec_kernel<DT_F16, 128><<<grid, ATT_THREADS, AttSpecCfg<128>::kSmemLaunch, st>>>(tq, tq, p);
      break;
    case 1:    // max-first kernel, P staged through shared memory
      if (bf) attention_fwd_kernel<DT_BF16, false><<<grid, ATT_THREADS, att_smem_bytes<false>(), st>>>(tq, p);
      else attention_fwd_kernel<DT_F16, false><<<grid, ATT_THREADS, att_smem_bytes<false>(), st>>>(tq, p);
      break;
    case 2:    // max-first kernel, P in TMEM (the round-1 baseline the default is measured against)
      if (bf) attention_fwd_kernel<DT_BF16, true><<<grid, ATT_THREADS, att_smem_bytes<true>(), st>>>(tq, p);
      else attention_fwd_kernel<DT_F16, true><<<grid, ATT_THREADS, att_smem_bytes<true>(), st>>>(tq, p);
      break;
    case 3:    // chains kernel, 3 chains x 128 keys (attention_chain.cuh); needs >= 2 KV tiles per item
    case 4:    // chains kernel, 4 chains x 96 keys
    case 5:    // chains kernel, 3 chains x 128 keys, two softmax warps per (chain, lane quadrant) splitting the columns
    case 7:    // chains kernel, 3 x 128, lean protocol
    case 8:    // chains kernel, 3 x 128, dedicated epilogue warpgroup
    case 9: {  // ... + PV starts on the first half of P
      const int bkv = variant == 4 ? AtcCfg4::BKV : AtcCfg3::BKV;
      const int nkv_max = variant == 4 ? AtcCfg4::NKV_MAX : variant == 5 ? AtcCfg3x2::NKV_MAX : AtcCfg3::NKV_MAX;
      if (N <= bkv || (N + bkv - 1) / bkv > nkv_max) return maest_attention_fwd(qkv, out, lse, B, N, H, op_dtype, 0, stream);
      const int items = B * H * ((N + ATT_BQ - 1) / ATT_BQ);
      const int sms = g_num_sms[cur_device()];
      const int g = items < sms ? items : sms;
      CUtensorMap tkv;
      if ((r = make_tmap(&tkv, qkv, op_dtype, uint64_t(B) * N, uint64_t(3) * H * ATT_D, uint64_t(3) * H * ATT_D, bkv / ATC_LOADDIV))) return r;
      if (variant == 3) {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg3><<<g, AtcCfg3::THREADS, AtcCfg3::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg3><<<g, AtcCfg3::THREADS, AtcCfg3::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      } else if (variant == 9) {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg3EH><<<g, AtcCfg3EH::THREADS, AtcCfg3EH::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg3EH><<<g, AtcCfg3EH::THREADS, AtcCfg3EH::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      } else if (variant == 8) {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg3E><<<g, AtcCfg3E::THREADS, AtcCfg3E::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg3E><<<g, AtcCfg3E::THREADS, AtcCfg3E::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      } else if (variant == 7) {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg3L><<<g, AtcCfg3L::THREADS, AtcCfg3L::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg3L><<<g, AtcCfg3L::THREADS, AtcCfg3L::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      } else if (variant == 5) {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg3x2><<<g, AtcCfg3x2::THREADS, AtcCfg3x2::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg3x2><<<g, AtcCfg3x2::THREADS, AtcCfg3x2::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      } else {
        if (bf) attention_fwd_chain_kernel<DT_BF16, AtcCfg4><<<g, AtcCfg4::THREADS, AtcCfg4::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
        else attention_fwd_chain_kernel<DT_F16, AtcCfg4><<<g, AtcCfg4::THREADS, AtcCfg4::SMEM_BYTES, st>>>(tq, tkv, p, qkv);
      }
      break;
    }
    case 16:   // timing diagnostic: the default kernel forced to ONE CTA per SM by padding the dynamic smem request
      if (bf) attention_fwd_spec_kernel<DT_BF16, 128><<<grid, ATT_THREADS, 120 * 1024, st>>>(tq, tq, p);
      else attention_fwd_spec_kernel<DT_F16, 128><<<grid, ATT_THREADS, 120 * 1024, st>>>(tq, tq, p);
      break;
    default: return fail(-1, "attention: unknown variant %d", variant);
  }
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

size_t maest_wave_tokens_workspace_bytes(int32_t B, int32_t S, int32_t P) {
  const size_t mel = (size_t(B) * LM_NMEL * size_t(1 + S / LM_HOP) * 4 + 255) & ~size_t(255);
  return mel + maest_patch_workspace_bytes(B, P);
}

int32_t maest_wave_tokens_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, const void* w_pe, int32_t op_dtype,
                              const float* conv_bias, const float* freq_pe, int32_t Fp, const float* time_pe, int32_t Wt,
                              const float* cls_token, const float* dist_token, const float* new_pos_embed, const int32_t* keep_ft,
                              int32_t P, int32_t t_offset, float* tokens, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return 0;
  if (workspace_bytes < maest_wave_tokens_workspace_bytes(B, S, P)) return fail(-1, "wave_tokens: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(-4, "wave_tokens: workspace must be 256-byte aligned");
  const int T = 1 + S / LM_HOP;
  const size_t mel_bytes = (size_t(B) * LM_NMEL * size_t(T) * 4 + 255) & ~size_t(255);
  float* mel = reinterpret_cast<float*>(workspace);
  int r;
  if ((r = logmel_launch(wav, B, S, wav_stride, mel, nullptr, stream))) return r;
  return maest_patch_tokens_fwd(mel, MAEST_F32, B, T, w_pe, op_dtype, conv_bias, freq_pe, Fp, time_pe, Wt, cls_token, dist_token,
                                new_pos_embed, keep_ft, P, t_offset, tokens, reinterpret_cast<uint8_t*>(workspace) + mel_bytes,
                                workspace_bytes - mel_bytes, stream);
}

size_t maest_encoder_workspace_bytes(int64_t rows) {
  // h16 [rows,768] | qkv16 [rows,2304] | o16 [rows,768] | u16 [rows,3072]  (+ 128 rows of slack per buffer) | LN stats [rows,2] fp32
  // | LN stats [rows,2] fp32 | LN partials [24,rows,4] fp32
  return size_t(rows + 128) * (768 + 2304 + 768 + 3072) * 2 + size_t(rows + 128) * (8 + 24 * 16) + 16 + 1024;
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
  float* stats = reinterpret_cast<float*>(u16 + R * 3072 * 2);
  float* parts = stats + ((R * 2 + 3) & ~size_t(3));   // 16-byte aligned records
  int r;
  // LayerNorm folding (23 of the 24 LayerNorms disappear into the GEMM epilogues around them) needs the folded vectors of
  // every block; block 0's norm1 reads the token buffer written by K2 and stays a kernel.
  bool fold = true;
  for (int i = 0; i < n_blocks; ++i) fold = fold && blocks[i].qkv_wg && blocks[i].qkv_bf && blocks[i].fc1_wg && blocks[i].fc1_bf;
  bool h16_is_folded = false;   // h16 holds x * gamma of the next LayerNorm and `stats` its row sums
  for (int i = 0; i < n_blocks; ++i) {
    const MaestBlockWeights& w = blocks[i];
    const bool attn_only = last_attn_only && i == n_blocks - 1;
    if (h16_is_folded) {
      if ((r = maest_linear_ln_fwd(h16, 768, w.qkv_w, 768, w.qkv_bf, int(M), 2304, 768, op_dtype, MAEST_EPI_STORE16_LN, qkv16, 2304,
                                   nullptr, stats, w.qkv_wg, nullptr, stream))) return r;
    } else {
      if ((r = maest_layernorm_fwd(x, w.ln1_w, w.ln1_b, h16, op_dtype, int(M), 1e-6f, nullptr, nullptr, stream))) return r;
      if ((r = maest_linear_fwd(h16, 768, w.qkv_w, 768, w.qkv_b, int(M), 2304, 768, op_dtype, MAEST_EPI_STORE16, qkv16, 2304,
                                nullptr, nullptr, 0, 0, 0, stream))) return r;
    }
    h16_is_folded = false;
    if ((r = maest_attention_fwd(qkv16, o16, nullptr, B, N, 12, op_dtype, attn_variant, stream))) return r;
    if (attn_only) {
      return maest_linear_fwd(o16, 768, w.proj_w, 768, w.proj_b, int(M), 768, 768, op_dtype, MAEST_EPI_STORE32, x, 768, nullptr,
                              nullptr, 0, 0, 0, stream);
    }
    if (fold) {
      // x += proj(o) and, in the same epilogue, h16 = x * gamma2 and the row statistics of x; fc1 finishes norm2
      if ((r = maest_linear_ln_fwd(o16, 768, w.proj_w, 768, w.proj_b, int(M), 768, 768, op_dtype, MAEST_EPI_RESID32_LN, x, 768, x,
                                   parts, w.ln2_w, h16, stream))) return r;
      if ((r = maest_ln_finalize(parts, int(M), 768, 1e-6f, stats, stream))) return r;
      if ((r = maest_linear_ln_fwd(h16, 768, w.fc1_w, 768, w.fc1_bf, int(M), 3072, 768, op_dtype, MAEST_EPI_GELU16_LN, u16, 3072,
                                   nullptr, stats, w.fc1_wg, nullptr, stream))) return r;
    } else {
      if ((r = maest_linear_fwd(o16, 768, w.proj_w, 768, w.proj_b, int(M), 768, 768, op_dtype, MAEST_EPI_RESID32, x, 768, x,
                                nullptr, 0, 0, 0, stream))) return r;
      if ((r = maest_layernorm_fwd(x, w.ln2_w, w.ln2_b, h16, op_dtype, int(M), 1e-6f, nullptr, nullptr, stream))) return r;
      if ((r = maest_linear_fwd(h16, 768, w.fc1_w, 768, w.fc1_b, int(M), 3072, 768, op_dtype, MAEST_EPI_GELU16, u16, 3072, nullptr,
                                nullptr, 0, 0, 0, stream))) return r;
    }
    if (fold && i + 1 < n_blocks) {
      // x += fc2(u) producing the operand and statistics of the NEXT block's norm1
      if ((r = maest_linear_ln_fwd(u16, 3072, w.fc2_w, 3072, w.fc2_b, int(M), 768, 3072, op_dtype, MAEST_EPI_RESID32_LN, x, 768, x,
                                   parts, blocks[i + 1].ln1_w, h16, stream))) return r;
      if ((r = maest_ln_finalize(parts, int(M), 768, 1e-6f, stats, stream))) return r;
      h16_is_folded = true;
    } else {
      if ((r = maest_linear_fwd(u16, 3072, w.fc2_w, 3072, w.fc2_b, int(M), 768, 3072, op_dtype, MAEST_EPI_RESID32, x, 768, x,
                                nullptr, 0, 0, 0, stream))) return r;
    }
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

// ------------------------------------------------------------------------------------------------ training step
int32_t maest_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta, float* dq32,
                            void* dqkv, int32_t B, int32_t N, int32_t H, int32_t op_dtype, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  if (H != 12) return fail(-1, "attention_bwd: H must be 12");
  if (reinterpret_cast<uintptr_t>(dqkv) & 31) return fail(-4, "attention_bwd: dqkv must be 32-byte aligned (256-bit stores)");
  cudaStream_t st = (cudaStream_t)stream;
  const long M = long(B) * N;
  CUtensorMap tq, td, tdq;
  int r;
  if ((r = make_tmap(&tq, qkv, op_dtype, M, 3 * H * 64, 3 * H * 64, 128))) return r;
  if ((r = make_tmap(&td, d_o, op_dtype, M, H * 64, H * 64, 128))) return r;
  if ((r = make_tmap_f32(&tdq, dq32, M, H * 64, H * 64, 128))) return r;
  CUDA_OK(cudaMemsetAsync(dq32, 0, size_t(M) * H * 64 * sizeof(float), st));
  AttnBwdParams p;
  p.B = B; p.N = N; p.H = H; p.lse = lse; p.delta = delta; p.dq32 = dq32; p.dqkv16 = dqkv;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  const long nd = M * H;
  const long n_items = long((N + 127) / 128) * H * B;          // persistent: one CTA per SM walks the (clip, head, key tile) items
  const int sms = g_num_sms[cur_device()];
  const unsigned grid = unsigned(n_items < sms ? n_items : sms);
  const int cast_blocks = int((M + 7) / 8);
  if (op_dtype == MAEST_BF16) {
    attn_delta_kernel<DT_BF16><<<unsigned((nd + 255) / 256), 256, 0, st>>>(o, d_o, delta, B, N, H);
    attention_bwd_kernel<DT_BF16><<<grid, ATTB_THREADS, ATTB_SMEM_BYTES, st>>>(tq, td, tdq, p);
    cast_rows16_kernel<DT_BF16><<<cast_blocks, 256, 0, st>>>(dq32, dqkv, 3 * H * 64, int(M), 0x7fffffff, 0, 0);
  } else if (op_dtype == MAEST_F16) {
    attn_delta_kernel<DT_F16><<<unsigned((nd + 255) / 256), 256, 0, st>>>(o, d_o, delta, B, N, H);
    attention_bwd_kernel<DT_F16><<<grid, ATTB_THREADS, ATTB_SMEM_BYTES, st>>>(tq, td, tdq, p);
    cast_rows16_kernel<DT_F16><<<cast_blocks, 256, 0, st>>>(dq32, dqkv, 3 * H * 64, int(M), 0x7fffffff, 0, 0);
  } else return fail(-1, "attention_bwd: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_mixup_fwd(const void* x, int32_t x_dtype, const int32_t* perm, const float* lam, float* out, int32_t B,
                        int64_t L, void* stream) {
  if (B <= 0 || L <= 0) return 0;
  dim3 grid(unsigned(L / 256 / 8 + 1 > 1024 ? 1024 : L / 256 / 8 + 1), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == MAEST_F16) mixup_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), perm, lam, out, B, L);
  else if (x_dtype == MAEST_F32) mixup_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), perm, lam, out, B, L);
  else return fail(-1, "mixup: input must be f16 or f32");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_bce_logits_fwd(const float* logits, const float* targets, int32_t n, float* loss, float* dlogits, void* stream) {
  if (n <= 0) return fail(-1, "bce: empty input");
  bce_logits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, targets, loss, dlogits, n);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_head_bwd(const float* x, int32_t B, int32_t N, const float* dlogits, const float* gscale, const float* norm_w,
                       const float* norm_b, const float* head_ln_w, const float* head_ln_b, const float* head_w, int32_t C,
                       float* dx, float* hz_ws, float* d_norm_w, float* d_norm_b, float* d_head_ln_w, float* d_head_ln_b,
                       float* d_head_w, float* d_head_b, void* stream) {
  if (B <= 0) return 0;
  if (C > 1024) return fail(-1, "head_bwd: at most 1024 classes");
  HeadBwdParams p;
  p.x = x; p.N = N; p.dlogits = dlogits; p.gscale = gscale; p.norm_w = norm_w; p.norm_b = norm_b; p.hln_w = head_ln_w;
  p.hln_b = head_ln_b; p.head_w = head_w; p.C = C; p.dx = dx; p.hz = hz_ws; p.d_norm_w = d_norm_w; p.d_norm_b = d_norm_b;
  p.d_hln_w = d_head_ln_w; p.d_hln_b = d_head_ln_b;
  p.separated = 0; p.dlogits_dist = nullptr; p.hdist_w = nullptr; p.z1out = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  head_bwd_clip_kernel<<<B, 256, 0, st>>>(p);
  head_wgrad_kernel<<<C, 256, 0, st>>>(dlogits, gscale, hz_ws, B, C, d_head_w, d_head_b);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_head_bwd_separated(const float* x, int32_t B, int32_t N, const float* dlogits, const float* dlogits_dist,
                                 const float* gscale, const float* norm_w, const float* norm_b, const float* head_ln_w,
                                 const float* head_ln_b, const float* head_w, const float* head_dist_w, int32_t C, float* dx,
                                 float* hz_ws, float* z1_ws, float* d_norm_w, float* d_norm_b, float* d_head_ln_w,
                                 float* d_head_ln_b, float* d_head_w, float* d_head_b, float* d_head_dist_w, float* d_head_dist_b,
                                 void* stream) {
  if (B <= 0) return 0;
  if (C > 1024) return fail(-1, "head_bwd_separated: at most 1024 classes");
  HeadBwdParams p;
  p.x = x; p.N = N; p.dlogits = dlogits; p.gscale = gscale; p.norm_w = norm_w; p.norm_b = norm_b; p.hln_w = head_ln_w;
  p.hln_b = head_ln_b; p.head_w = head_w; p.C = C; p.dx = dx; p.hz = hz_ws; p.d_norm_w = d_norm_w; p.d_norm_b = d_norm_b;
  p.d_hln_w = d_head_ln_w; p.d_hln_b = d_head_ln_b;
  p.separated = 1; p.dlogits_dist = dlogits_dist; p.hdist_w = head_dist_w; p.z1out = z1_ws;
  cudaStream_t st = (cudaStream_t)stream;
  head_bwd_clip_kernel<<<B, 256, 0, st>>>(p);
  head_wgrad_kernel<<<C, 256, 0, st>>>(dlogits, gscale, hz_ws, B, C, d_head_w, d_head_b);
  head_wgrad_kernel<<<C, 256, 0, st>>>(dlogits_dist, gscale, z1_ws, B, C, d_head_dist_w, d_head_dist_b);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                            float* dx, void* dx16, int32_t op_dtype, float* dgamma, float* dbeta, float* dx_colsum, int32_t rows,
                            void* stream) {
  if (rows <= 0) return 0;
  const int sms = g_num_sms[cur_device()] > 0 ? g_num_sms[cur_device()] : 148;
  int blocks = (rows + 7) / 8;
  if (blocks > sms * 4) blocks = sms * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == MAEST_BF16) layernorm_bwd_kernel<DT_BF16><<<blocks, 256, 0, st>>>(dy, x, mean, rstd, gamma, dx, dx16, dgamma, dbeta, dx_colsum, rows);
  else if (op_dtype == MAEST_F16) layernorm_bwd_kernel<DT_F16><<<blocks, 256, 0, st>>>(dy, x, mean, rstd, gamma, dx, dx16, dgamma, dbeta, dx_colsum, rows);
  else return fail(-1, "layernorm_bwd: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_colsum(const void* in, int32_t in_dtype, int64_t ld, int32_t M, int32_t N, float* out, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (N % 8 || ld % 8) return fail(-1, "colsum: N and ld must be multiples of 8");
  int splits = (M + 255) / 256;
  splits = splits < 1 ? 1 : (splits > 64 ? 64 : splits);
  dim3 grid((N + 255) / 256, splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == MAEST_F32) colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), ld, M, N, out);
  else if (in_dtype == MAEST_F16) colsum_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(in), ld, M, N, out);
  else if (in_dtype == MAEST_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld, M, N, out);
  else return fail(-1, "colsum: bad dtype");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_cast_rows16(const float* src, void* dst, int64_t dst_ld, int32_t rows, int32_t rows_per_group, int32_t group_stride,
                          int32_t row_offset, int32_t op_dtype, void* stream) {
  if (rows <= 0) return 0;
  if (rows_per_group <= 0) { rows_per_group = 0x7fffffff; group_stride = 0; row_offset = 0; }
  const int blocks = (rows + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == MAEST_BF16) cast_rows16_kernel<DT_BF16><<<blocks, 256, 0, st>>>(src, dst, dst_ld, rows, rows_per_group, group_stride, row_offset);
  else if (op_dtype == MAEST_F16) cast_rows16_kernel<DT_F16><<<blocks, 256, 0, st>>>(src, dst, dst_ld, rows, rows_per_group, group_stride, row_offset);
  else return fail(-1, "cast_rows16: op_dtype must be f16/bf16");
  CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t maest_token_grad(const float* dx, int32_t B, int32_t N, int32_t P, int32_t Tp, int32_t Fp, int32_t Wt, int32_t t_offset,
                         const int32_t* keep_ft, float* d_cls, float* d_dist, float* d_new_pos, float* d_conv_bias,
                         float* d_freq, float* d_time, void* stream) {
  if (B <= 0) return 0;
  TokenGradParams p;
  p.dx = dx; p.B = B; p.N = N; p.P = P; p.Tp = Tp; p.Fp = Fp; p.Wt = Wt; p.t_off = t_offset; p.keep_ft = keep_ft;
  p.d_cls = d_cls; p.d_dist = d_dist; p.d_new_pos = d_new_pos; p.d_conv_bias = d_conv_bias; p.d_freq = d_freq; p.d_time = d_time;
  token_grad_kernel<<<2 + P, 256, 0, (cudaStream_t)stream>>>(p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
