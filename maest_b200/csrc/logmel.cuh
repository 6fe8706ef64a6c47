// K1: waveform -> normalised log-mel, one fused kernel (fp32 throughout).
//
// Replaces models/helpers/melspectrogram.py:47-60 and the torchaudio/cuFFT/cuBLAS chain under it
// (reflect pad, framing x Hann, 512-pt rFFT, |.|^2, [T,257]x[257,96] mel matmul, log10(1+1e4 x), z-norm):
// ~10 library kernels with a [B,257,T] HBM round trip each become one pass: waveform in, mel out.
//
// Work decomposition: CTA = (clip, chunk of LM_FRAMES consecutive frames), 256 threads = 4 groups of 64.
// A group transforms TWO real frames with ONE 512-point complex FFT (frame A -> real, frame B -> imag),
// done as three radix-8 passes (512 = 8*8*8) with the 8-point DFTs in registers and two conflict-free
// shared-memory exchanges; spectra are separated with the conjugate-symmetry identity, the mel filterbank
// is applied in its sparse triangular form (<= 15 bins per band, 502 non-zeros), and results are staged
// in smem so that global stores are contiguous along time.
//
// The per-thread phase functions below are __host__ __device__: tests/test_api_cpu.py::test_logmel_kernel_host_emulation
// compiles this header with g++ (tests/logmel_host_emu.cpp, MB_HOST_EMULATION) and replays the exact index arithmetic on the CPU.
#pragma once
#ifndef MB_HOST_EMULATION
#include "common.cuh"
#define MB_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <stdint.h>
#define MB_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif

namespace mb {

constexpr int LM_NFFT = 512;
constexpr int LM_HOP = 256;
constexpr int LM_NMEL = 96;
// 16 frames per CTA and three CTAs per SM (74 KB of shared memory each): 0.198 ms against 0.214 ms for 32 frames / two CTAs
// at config 3 (the kernel is latency-bound on its shared-memory exchanges: 24 resident warps hide more of it than 16).
#ifndef LM_FRAMES_PER_CTA
#define LM_FRAMES_PER_CTA 16
#define LM_CTAS_PER_SM 3
#endif
constexpr int LM_FRAMES = LM_FRAMES_PER_CTA;  // frames per CTA
constexpr int LM_GROUPS = 4;                  // FFT groups per CTA (64 threads each)
constexpr int LM_SEG = (LM_FRAMES + 1) * LM_HOP;   // waveform samples staged per CTA
constexpr int LM_A_STRIDE = 72;               // pass-1 -> pass-2 exchange: [k0][n1*8+n0], row stride 72
constexpr int LM_B_STRIDE = 65;               // pass-2 -> pass-3 exchange: [n0][k0*8+k1], row stride 65
constexpr int LM_MAX_TAPS = 16;

// tables shared by all CTAs (built on the host in double precision)
struct LogMelTables {
  float2 tw[LM_NFFT];        // W512^j = exp(-2 pi i j / 512)
  float hann[LM_NFFT];       // periodic Hann (torchaudio framing); replaced in smem by hann_sym for the Essentia framing
  int band_start[LM_NMEL];   // first FFT bin with non-zero weight
  int band_len[LM_NMEL];     // number of bins (<= LM_MAX_TAPS)
  float band_w[LM_NMEL * LM_MAX_TAPS];
  float hann_sym[LM_NFFT];   // symmetric Hann 0.5 - 0.5 cos(2 pi j / 511): Essentia's Windowing(type='hann')
};

// per-group scratch (floats): bufA re/im [8*72], bufB re/im [8*65]
constexpr int LM_BUFA = 8 * LM_A_STRIDE;      // 576
constexpr int LM_BUFB = 8 * LM_B_STRIDE;      // 520
constexpr int LM_GROUP_FLOATS = 2 * LM_BUFA + 2 * LM_BUFB;

MB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place forward 8-point DFT, natural order in and out
MB_HD void dft8(float2 (&v)[8]) {
  const float r = 0.70710678118654752f;
  float2 a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = make_float2(v[j].x + v[j + 4].x, v[j].y + v[j + 4].y);
    b[j] = make_float2(v[j].x - v[j + 4].x, v[j].y - v[j + 4].y);
  }
  // b[j] *= W8^j
  b[1] = make_float2((b[1].x + b[1].y) * r, (b[1].y - b[1].x) * r);
  b[2] = make_float2(b[2].y, -b[2].x);
  b[3] = make_float2((b[3].y - b[3].x) * r, -(b[3].x + b[3].y) * r);
  // DFT4(a) -> X[0,2,4,6];  DFT4(b) -> X[1,3,5,7]
  {
    float2 e0 = make_float2(a[0].x + a[2].x, a[0].y + a[2].y), e1 = make_float2(a[0].x - a[2].x, a[0].y - a[2].y);
    float2 o0 = make_float2(a[1].x + a[3].x, a[1].y + a[3].y), d = make_float2(a[1].x - a[3].x, a[1].y - a[3].y);
    float2 o1 = make_float2(d.y, -d.x);
    v[0] = make_float2(e0.x + o0.x, e0.y + o0.y);
    v[4] = make_float2(e0.x - o0.x, e0.y - o0.y);
    v[2] = make_float2(e1.x + o1.x, e1.y + o1.y);
    v[6] = make_float2(e1.x - o1.x, e1.y - o1.y);
  }
  {
    float2 e0 = make_float2(b[0].x + b[2].x, b[0].y + b[2].y), e1 = make_float2(b[0].x - b[2].x, b[0].y - b[2].y);
    float2 o0 = make_float2(b[1].x + b[3].x, b[1].y + b[3].y), d = make_float2(b[1].x - b[3].x, b[1].y - b[3].y);
    float2 o1 = make_float2(d.y, -d.x);
    v[1] = make_float2(e0.x + o0.x, e0.y + o0.y);
    v[5] = make_float2(e0.x - o0.x, e0.y - o0.y);
    v[3] = make_float2(e1.x + o1.x, e1.y + o1.y);
    v[7] = make_float2(e1.x - o1.x, e1.y - o1.y);
  }
}

// Index algebra: n = 64 n2 + 8 n1 + n0, k = k0 + 8 k1 + 64 k2,
//   n k = 64 n2 k0 + 8 n1 k0 + 64 n1 k1 + n0 (k0 + 8 k1) + 64 n0 k2   (mod 512)
// pass 1: DFT8 over n2, twiddle W512^(8 n1 k0);  pass 2: DFT8 over n1, twiddle W512^(n0 (k0 + 8 k1));
// pass 3: DFT8 over n0 -> Z[k0 + 8 k1 + 64 k2].

// pass 1, thread tid = 8 n1 + n0.  segA/segB: the two frames' 512 samples (already reflect-resolved).
MB_HD void lm_pass1(int tid, const float* segA, const float* segB, bool haveB, const LogMelTables& tb,
                    float* bufA_re, float* bufA_im) {
  float2 v[8];
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) {
    const int n = 64 * n2 + tid;
    const float w = tb.hann[n];
    v[n2] = make_float2(segA[n] * w, haveB ? segB[n] * w : 0.f);
  }
  dft8(v);
  const int n1 = tid >> 3;
#pragma unroll
  for (int k0 = 0; k0 < 8; ++k0) {
    const float2 t = cmul(v[k0], tb.tw[(8 * n1 * k0) & 511]);
    bufA_re[k0 * LM_A_STRIDE + tid] = t.x;
    bufA_im[k0 * LM_A_STRIDE + tid] = t.y;
  }
}

// pass 2, thread tid = 8 k0 + n0
MB_HD void lm_pass2(int tid, const LogMelTables& tb, const float* bufA_re, const float* bufA_im, float* bufB_re,
                    float* bufB_im) {
  const int k0 = tid >> 3, n0 = tid & 7;
  float2 v[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1)
    v[n1] = make_float2(bufA_re[k0 * LM_A_STRIDE + n1 * 8 + n0], bufA_im[k0 * LM_A_STRIDE + n1 * 8 + n0]);
  dft8(v);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    const float2 t = cmul(v[k1], tb.tw[(n0 * (k0 + 8 * k1)) & 511]);
    bufB_re[n0 * LM_B_STRIDE + k0 * 8 + k1] = t.x;
    bufB_im[n0 * LM_B_STRIDE + k0 * 8 + k1] = t.y;
  }
}

// pass 3, thread tid = 8 k0 + k1: writes Z[k] (k = k0 + 8 k1 + 64 k2) into z_re/z_im[512] (aliases bufA)
MB_HD void lm_pass3(int tid, const float* bufB_re, const float* bufB_im, float* z_re, float* z_im) {
  const int k0 = tid >> 3, k1 = tid & 7;
  float2 v[8];
#pragma unroll
  for (int n0 = 0; n0 < 8; ++n0) v[n0] = make_float2(bufB_re[n0 * LM_B_STRIDE + tid], bufB_im[n0 * LM_B_STRIDE + tid]);
  dft8(v);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const int k = k0 + 8 * k1 + 64 * k2;
    z_re[k] = v[k2].x;
    z_im[k] = v[k2].y;
  }
}

// power spectra of the two real frames from Z = FFT(a + i b):
//   A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2 i);  bins 1..255 (bins 0 and 256 carry
//   zero filterbank weight).  Thread tid handles bins tid, tid+64, tid+128, tid+192.
MB_HD void lm_power(int tid, const float* z_re, const float* z_im, float* powA, float* powB) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = tid + 64 * i;
    if (k == 0) { powA[0] = 0.f; powB[0] = 0.f; continue; }
    const float zr = z_re[k], zi = z_im[k], yr = z_re[512 - k], yi = z_im[512 - k];
    const float ar = 0.5f * (zr + yr), ai = 0.5f * (zi - yi);
    const float br = 0.5f * (zi + yi), bi = 0.5f * (yr - zr);
    powA[k] = ar * ar + ai * ai;
    powB[k] = br * br + bi * bi;
  }
}

// sparse triangular filterbank + log compression for one (band, frame): log10(1 + 1e4 x) = log2(.) * log10(2)
MB_HD float lm_band_raw(int band, const float* pw, const LogMelTables& tb) {
  const int s = tb.band_start[band], n = tb.band_len[band];
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc = fmaf(pw[s + i], tb.band_w[band * LM_MAX_TAPS + i], acc);
  return log2f(fmaf(acc, 10000.0f, 1.0f)) * 0.30102999566398120f;
}
// z-norm of the model input: (x - 2.06755686098554) / (2 * 1.268292820667291)   (models/helpers/melspectrogram.py:57-60)
MB_HD float lm_znorm(float lg) { return (lg - 2.06755686098554f) * 0.39423072641610746f; }
MB_HD float lm_band(int band, const float* pw, const LogMelTables& tb) { return lm_znorm(lm_band_raw(band, pw, tb)); }

MB_HD int lm_reflect(int s, int S) {
  if (s < 0) s = -s;
  if (s >= S) s = 2 * (S - 1) - s;
  return s;
}

#ifndef MB_HOST_EMULATION
struct LogMelParams {
  const float* wav;   // [B, wav_stride]
  long wav_stride;
  int B, S, T;
  float* mel;         // [B, 96, T] normalised fp32 (model input), or null
  __half* raw_tm16;   // [B, T, 96] un-normalised log10(1 + 1e4 mel) as fp16, time-major: the layout of the reference's
                      // .mmap training files (helpers/melspectrogram_extractor.py:45-47), or null
  const LogMelTables* tables;
  int essentia_framing;   // 0: torchaudio (reflect padding, periodic Hann, T = 1 + S / 256) -- the model's front-end;
                          // 1: Essentia FrameCutter(startFromZero=false) + Windowing('hann', normalized=false): frames centred on
                          //    256 t, ZERO padding outside the signal, symmetric Hann, T = ceil(S / 256) -- what the offline
                          //    extractor of the reference runs (helpers/melspectrogram_extractor.py:15-30)
};

constexpr int LM_THREADS = 64 * LM_GROUPS;
// smem: tables 14.8 KB | waveform segment 17 KB | group scratch 4 x 8.6 KB | out staging 96 x 17 x 4 = 6.4 KB  (74 KB: 3 CTAs / SM)
constexpr int LM_OUT_STRIDE = LM_FRAMES + 1;
constexpr int LM_SMEM_BYTES = int(sizeof(LogMelTables)) + LM_SEG * 4 + LM_GROUPS * LM_GROUP_FLOATS * 4 +
                              LM_NMEL * LM_OUT_STRIDE * 4;

__global__ void __launch_bounds__(LM_THREADS, LM_CTAS_PER_SM) logmel_kernel(const LogMelParams p) {
  extern __shared__ __align__(16) uint8_t lm_smem[];
  LogMelTables& tb = *reinterpret_cast<LogMelTables*>(lm_smem);
  float* seg = reinterpret_cast<float*>(lm_smem + sizeof(LogMelTables));
  float* scratch = seg + LM_SEG;
  float* outs = scratch + LM_GROUPS * LM_GROUP_FLOATS;

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * LM_FRAMES;
  const int nfr = min(LM_FRAMES, p.T - t0);
  const int tid = threadIdx.x;

  // Staging through the bulk-copy (TMA) engine: the 15 KB of tables and -- for every CTA whose segment lies inside the signal,
  // i.e. all but the first and last of a clip -- the 33 KB waveform segment arrive as two cp.async.bulk transfers issued by one
  // thread and tracked by one mbarrier (no per-thread load / store instructions, no registers in between).  Edge CTAs resolve
  // the reflect / zero padding element by element as before.
  __shared__ __align__(8) uint64_t lm_bar;
  const float* x = p.wav + long(b) * p.wav_stride;
  const int base = LM_HOP * (t0 - 1);
  const int need = (nfr + 1) * LM_HOP;
  const bool bulk_seg = base >= 0 && base + need <= p.S && ((reinterpret_cast<uintptr_t>(x + base) & 15) == 0);
  if (tid == 0) {
    mbar_init(&lm_bar, 1);
    fence_mbar_init();
    const uint32_t bytes = uint32_t(sizeof(LogMelTables)) + (bulk_seg ? uint32_t(need) * 4u : 0u);
    mbar_expect_tx(&lm_bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(lm_smem)),
                 "l"(p.tables), "r"(uint32_t(sizeof(LogMelTables))), "r"(smem_u32(&lm_bar)) : "memory");
    if (bulk_seg)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(seg)),
                   "l"(x + base), "r"(uint32_t(need) * 4u), "r"(smem_u32(&lm_bar)) : "memory");
  }
  if (!bulk_seg) {  // waveform segment with the padding resolved: seg[i] = x[reflect(256 (t0 - 1) + i)] (or 0 outside the signal)
    if (p.essentia_framing) {
      for (int i = tid; i < need; i += LM_THREADS) {
        const int s = base + i;
        seg[i] = (s >= 0 && s < p.S) ? __ldg(x + s) : 0.f;
      }
    } else {
      for (int i = tid; i < need; i += LM_THREADS) seg[i] = __ldg(x + lm_reflect(base + i, p.S));
    }
  }
  __syncthreads();               // lm_bar initialised (and the edge CTAs' segment stores done) before anyone waits
  mbar_wait(&lm_bar, 0);
  if (p.essentia_framing) {
    for (int i = tid; i < LM_NFFT; i += LM_THREADS) tb.hann[i] = tb.hann_sym[i];
    __syncthreads();
  }

  const int g = tid >> 6, gt = tid & 63;
  float* bufA_re = scratch + g * LM_GROUP_FLOATS;
  float* bufA_im = bufA_re + LM_BUFA;
  float* bufB_re = bufA_im + LM_BUFA;
  float* bufB_im = bufB_re + LM_BUFB;
  const int bar_id = 1 + g;  // named barrier per 64-thread group
  auto gsync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };

  const int npairs = (nfr + 1) >> 1;
  for (int pr = g; pr < npairs; pr += LM_GROUPS) {
    const int fa = 2 * pr, fb = fa + 1;
    const bool haveB = fb < nfr;
    lm_pass1(gt, seg + fa * LM_HOP, seg + fb * LM_HOP, haveB, tb, bufA_re, bufA_im);
    gsync();
    lm_pass2(gt, tb, bufA_re, bufA_im, bufB_re, bufB_im);
    gsync();
    lm_pass3(gt, bufB_re, bufB_im, bufA_re, bufA_im);   // Z re -> bufA_re[0..511], Z im -> bufA_im[0..511]
    gsync();
    lm_power(gt, bufA_re, bufA_im, bufB_re, bufB_re + 256);  // powA -> bufB_re[0..255], powB -> [256..511]
    gsync();
    for (int o = gt; o < 2 * LM_NMEL; o += 64) {
      const int which = o / LM_NMEL, band = o - which * LM_NMEL;
      if (which == 0 || haveB) outs[band * LM_OUT_STRIDE + fa + which] = lm_band_raw(band, bufB_re + which * 256, tb);
    }
    gsync();
  }
  __syncthreads();
  if (p.mel != nullptr) {   // coalesced store: consecutive threads -> consecutive frames of one band
    float* mel = p.mel + long(b) * LM_NMEL * p.T;
    for (int i = tid; i < LM_NMEL * LM_FRAMES; i += LM_THREADS) {
      const int band = i / LM_FRAMES, f = i - band * LM_FRAMES;
      if (f < nfr) mel[long(band) * p.T + t0 + f] = lm_znorm(outs[band * LM_OUT_STRIDE + f]);
    }
  }
  if (p.raw_tm16 != nullptr) {   // time-major: consecutive threads -> consecutive bands of one frame (192-byte rows)
    __half* raw = p.raw_tm16 + (long(b) * p.T + t0) * LM_NMEL;
    for (int i = tid; i < LM_NMEL * LM_FRAMES; i += LM_THREADS) {
      const int f = i / LM_NMEL, band = i - f * LM_NMEL;
      if (f < nfr) raw[long(f) * LM_NMEL + band] = __float2half_rn(outs[band * LM_OUT_STRIDE + f]);
    }
  }
}
#endif  // !MB_HOST_EMULATION

}  // namespace mb
