// Host-side construction of the K1 tables (double precision, rounded once to fp32).
// Filterbank: Slaney mel scale + Slaney area normalisation, 96 bands over bins linspace(0, 8000, 257),
// i.e. what models/helpers/melspectrogram.py:36-42 asks torchaudio's MelScale for.
#pragma once
#include <math.h>
#include <string.h>

#include "logmel.cuh"

namespace mb {

inline double lm_hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
inline double lm_mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// returns 0 on success, -1 if a band needs more than LM_MAX_TAPS bins
inline int build_logmel_tables(LogMelTables* tb) {
  memset(tb, 0, sizeof(*tb));
  const double PI = 3.14159265358979323846;
  for (int j = 0; j < LM_NFFT; ++j) {
    const double a = -2.0 * PI * j / LM_NFFT;
    tb->tw[j].x = (float)cos(a);
    tb->tw[j].y = (float)sin(a);
    tb->hann[j] = (float)(0.5 - 0.5 * cos(2.0 * PI * j / LM_NFFT));
    tb->hann_sym[j] = (float)(0.5 - 0.5 * cos(2.0 * PI * j / (LM_NFFT - 1)));
  }
  const int n_freqs = LM_NFFT / 2 + 1;
  double f_pts[LM_NMEL + 2];
  const double m_min = lm_hz_to_mel(0.0), m_max = lm_hz_to_mel(8000.0);
  for (int i = 0; i < LM_NMEL + 2; ++i) f_pts[i] = lm_mel_to_hz(m_min + (m_max - m_min) * i / (LM_NMEL + 1));
  for (int b = 0; b < LM_NMEL; ++b) {
    const double lo = f_pts[b], ce = f_pts[b + 1], hi = f_pts[b + 2];
    const double enorm = 2.0 / (hi - lo);
    int start = -1, len = 0;
    for (int k = 0; k < n_freqs; ++k) {
      const double f = 8000.0 * k / (n_freqs - 1);
      const double down = (f - lo) / (ce - lo), up = (hi - f) / (hi - ce);
      double w = down < up ? down : up;
      if (w <= 0.0) continue;
      if (start < 0) start = k;
      if (k - start >= LM_MAX_TAPS) return -1;
      tb->band_w[b * LM_MAX_TAPS + (k - start)] = (float)(w * enorm);
      len = k - start + 1;
    }
    if (start < 0) { start = 1; len = 0; }
    tb->band_start[b] = start;
    tb->band_len[b] = len;
  }
  return 0;
}

}  // namespace mb
