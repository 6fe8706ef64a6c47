// CPU replay of the K1 (log-mel) kernel's per-thread phase functions: validates the radix-8x8x8 FFT index
// algebra, the two-frames-per-FFT separation, the sparse filterbank and the reflect padding without a GPU.
// Built by tests/test_logmel_host_emulation.py:  g++ -O2 -DMB_HOST_EMULATION -I maest_b200/csrc ...
// usage: logmel_host_emu S < wave.f32 > mel.f32   (mel layout [96, T])
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "logmel_tables.h"

using namespace mb;

int main(int argc, char** argv) {
  const int S = atoi(argv[1]);
  const int T = 1 + S / LM_HOP;
  std::vector<float> x(S);
  if (fread(x.data(), 4, S, stdin) != (size_t)S) return 2;
  static LogMelTables tb;
  if (build_logmel_tables(&tb)) return 3;
  std::vector<float> mel((size_t)LM_NMEL * T);
  std::vector<float> seg(LM_SEG);
  float bufA_re[LM_BUFA], bufA_im[LM_BUFA], bufB_re[LM_BUFB], bufB_im[LM_BUFB];
  for (int t0 = 0; t0 < T; t0 += LM_FRAMES) {
    const int nfr = (T - t0) < LM_FRAMES ? (T - t0) : LM_FRAMES;
    const int base = LM_HOP * (t0 - 1);
    for (int i = 0; i < (nfr + 1) * LM_HOP; ++i) seg[i] = x[lm_reflect(base + i, S)];
    for (int pr = 0; pr < (nfr + 1) / 2; ++pr) {
      const int fa = 2 * pr, fb = fa + 1;
      const bool haveB = fb < nfr;
      for (int t = 0; t < 64; ++t) lm_pass1(t, &seg[fa * LM_HOP], &seg[fb * LM_HOP], haveB, tb, bufA_re, bufA_im);
      for (int t = 0; t < 64; ++t) lm_pass2(t, tb, bufA_re, bufA_im, bufB_re, bufB_im);
      for (int t = 0; t < 64; ++t) lm_pass3(t, bufB_re, bufB_im, bufA_re, bufA_im);
      for (int t = 0; t < 64; ++t) lm_power(t, bufA_re, bufA_im, bufB_re, bufB_re + 256);
      for (int o = 0; o < 2 * LM_NMEL; ++o) {
        const int which = o / LM_NMEL, band = o - which * LM_NMEL;
        if (which == 0 || haveB) mel[(size_t)band * T + t0 + fa + which] = lm_band(band, bufB_re + which * 256, tb);
      }
    }
  }
  fwrite(mel.data(), 4, mel.size(), stdout);
  return 0;
}
