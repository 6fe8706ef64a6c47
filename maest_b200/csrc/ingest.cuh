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

constexpr int ING_FRAMES = 64;   // output frames per CTA

__global__ void __launch_bounds__(256) mel_ingest_kernel(const IngestParams p) {
  __shared__ __half tile[ING_FRAMES][98];   // [frame][band], padded
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * ING_FRAMES;
  const int nfr = min(ING_FRAMES, p.T - t0);
  const int n = p.frames_read ? p.frames_read[b] : p.T;
  const int pad_half = (p.T - n) / 2;       // np.roll(padded, padding_size // 2, axis=0)
  int sf = p.roll_shift ? p.roll_shift[b] % p.T : 0;
  if (sf < 0) sf += p.T;
  const __half2* raw2 = reinterpret_cast<const __half2*>(p.raw + long(b) * p.T * 96);
  const __half2 mean2 = __half2half2(p.norm_mean), std2 = __half2half2(p.norm_2std);
  for (int i = threadIdx.x; i < nfr * 48; i += 256) {
    const int f = i / 48, bp = i - f * 48;
    int src = t0 + f - sf;                  // undo the time roll
    if (src < 0) src += p.T;
    src -= pad_half;                        // undo the centring roll of the zero padding
    if (src < 0) src += p.T;
    __half2 v = src < n ? raw2[long(src) * 48 + bp] : __half2half2(__ushort_as_half(0));
    if (p.do_norm) v = __h2div(__hsub2(v, mean2), std2);
    tile[f][2 * bp] = __low2half(v);
    tile[f][2 * bp + 1] = __high2half(v);
  }
  __syncthreads();
  __half* out = p.out + long(b) * 96 * p.T + t0;
  for (int i = threadIdx.x; i < 96 * ING_FRAMES; i += 256) {
    const int band = i / ING_FRAMES, f = i - band * ING_FRAMES;
    if (f < nfr) out[long(band) * p.T + f] = tile[f][band];
  }
}

}  // namespace mb
