// Loader -> device ingest: the reference reads a window of a raw float16 [frames, 96]
// .mmap file per clip on the host, zero-pads short windows and centres the padding with np.roll, transposes to
// [1, 96, T] (discogs/dataset.py:88-139), then normalises (discogs/datamodule.py:126-137, float16 arithmetic) and
// optionally rolls along time (datamodule.py:111-123) -- four numpy/torch passes per clip in the DataLoader workers.
// Here the raw window bytes go to the GPU as they lie in the file (time-major) and ONE kernel does all of it,
// writing the band-major fp16 batch [B, 1, 96, T] that K2 / the training step consume.
//
// Bit-exactness: the reference subtracts / divides float16 arrays by Python floats, i.e. numpy rounds the constants to
// float16 and rounds after each operation; __hsub / __hdiv round the same way (fp32 intermediate, 24 >= 2*11+2 bits).
#pragma once
#include "common.cuh"

namespace mb {

struct IngestParams {
  const __half* raw;        // [B, T, 96] time-major windows; rows >= frames_read[b] are ignored (may be garbage)
  const int32_t* frames_read;  // [B] frames actually read from the file (<= T), or null = T for every clip
  const int32_t* roll_shift;   // [B] time roll (torch.roll semantics: out[t] = in[(t - shift) mod T]), or null
  int B, T;
  int do_norm;
  __half norm_mean, norm_2std;
  __half* out;              // [B, 96, T]
};

constexpr int ING_FRAMES = 64;    // output frames per CTA
constexpr int ING_THREADS = 192;  // load phase: 48 band pairs x 4 frames per pass; store phase: 32 frame pairs x 6 bands per pass

__global__ void __launch_bounds__(ING_THREADS) mel_ingest_kernel(const IngestParams p) {
  __shared__ __align__(4) __half tile[96][ING_FRAMES + 2];   // [band][frame]: rows of 33 words, conflict-free in both phases
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * ING_FRAMES;
  const int nfr = min(ING_FRAMES, p.T - t0);
  const int n = p.frames_read ? p.frames_read[b] : p.T;
  const int pad_half = (p.T - n) / 2;       // np.roll(padded, padding_size // 2, axis=0)
  int sf = p.roll_shift ? p.roll_shift[b] % p.T : 0;
  if (sf < 0) sf += p.T;
  const __half2* raw2 = reinterpret_cast<const __half2*>(p.raw + long(b) * p.T * 96);
  const __half2 mean2 = __half2half2(p.norm_mean), std2 = __half2half2(p.norm_2std);
  {   // load: consecutive threads -> consecutive band pairs of one frame (192-byte rows of the file window)
    const int bp = threadIdx.x % 48, f0 = threadIdx.x / 48;
    int src = t0 + f0 - sf;                 // undo the time roll ...
    if (src < 0) src += p.T;
    src -= pad_half;                        // ... and the centring roll of the zero padding
    if (src < 0) src += p.T;
    for (int f = f0; f < nfr; f += 4) {
      __half2 v = src < n ? __ldg(raw2 + long(src) * 48 + bp) : __half2half2(__ushort_as_half(0));
      if (p.do_norm) v = __h2div(__hsub2(v, mean2), std2);
      tile[2 * bp][f] = __low2half(v);
      tile[2 * bp + 1][f] = __high2half(v);
      src += 4;
      if (src >= p.T) src -= p.T;
    }
  }
  __syncthreads();
  {   // store: consecutive threads -> consecutive frame PAIRS of one band; 4-byte stores wherever the row allows it (rows
      // of T fp16 start on odd element offsets for every other band when T is odd)
    const int k = threadIdx.x % 32, band0 = threadIdx.x / 32;
    for (int band = band0; band < 96; band += 6) {
      const long row = (long(b) * 96 + band) * p.T + t0;
      __half* out = p.out + row;
      const int mis = int(row & 1);          // 1: element 0 of this tile row is not 4-byte aligned
      const int f = 2 * k + mis;             // this thread's pair covers frames f, f + 1
      if (f + 1 < nfr) {
        const __half2 v = __halves2half2(tile[band][f], tile[band][f + 1]);
        *reinterpret_cast<__half2*>(out + f) = v;
      } else if (f < nfr) {
        out[f] = tile[band][f];
      }
      if (mis && k == 0 && nfr > 0) out[0] = tile[band][0];
    }
  }
}

}  // namespace mb
