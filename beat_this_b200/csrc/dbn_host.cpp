// Host Viterbi decoder of the bar-pointer HMM used by the DBN post-processor (beat_this_b200/dbn.py; stand-in
// for madmom's Cython HMM.viterbi behind DBNDownBeatTrackingProcessor, reference model/postprocessor.py:29-37,170).
// The state space is never materialised as a transition matrix: inside a beat a state can only be reached from
// the previous position of the same tempo, and the first position of a beat from the LAST position of every
// tempo of the previous beat (n_int x n_int log-probabilities).  Plain C++, no CUDA: called per clip from a
// Python thread pool (ctypes releases the GIL).
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../include/beatthis.h"

extern "C" int bt_dbn_viterbi(const double* log_dens, int64_t T, int32_t beats, int32_t n_int, const int32_t* intervals,
                              const double* log_tempo, const int32_t* pointers, int64_t* path_out, double* logp_out) {
  if (!log_dens || !intervals || !log_tempo || !pointers || !path_out || !logp_out || T <= 0 || beats <= 0 || n_int <= 0)
    return BT_ERR_ARG;
  int64_t per_beat = 0;
  std::vector<int64_t> first(n_int), last(n_int);
  for (int k = 0; k < n_int; ++k) {
    if (intervals[k] <= 0) return BT_ERR_ARG;
    first[k] = per_beat;
    per_beat += intervals[k];
    last[k] = per_beat - 1;
  }
  const int64_t S = per_beat * beats;
  std::vector<double> v(S, -std::log(static_cast<double>(S))), nv(S);
  std::vector<int16_t> back(static_cast<size_t>(T) * beats * n_int);
  std::vector<double> from(n_int);
  const double ninf = -std::numeric_limits<double>::infinity();
  for (int64_t t = 0; t < T; ++t) {
    const double* d = log_dens + 3 * t;
    for (int b = 0; b < beats; ++b) {
      const int64_t base = b * per_beat, prev_base = ((b + beats - 1) % beats) * per_beat;
      for (int k = 0; k < n_int; ++k) from[k] = v[prev_base + last[k]];
      for (int k = 0; k < n_int; ++k) {  // first position of tempo k: best previous tempo
        double best = ninf;
        int arg = 0;
        for (int f = 0; f < n_int; ++f) {
          const double c = from[f] + log_tempo[f * n_int + k];
          if (c > best) { best = c; arg = f; }
        }
        back[(static_cast<size_t>(t) * beats + b) * n_int + k] = static_cast<int16_t>(arg);
        const int64_t s0 = base + first[k];
        nv[s0] = best + d[pointers[s0]];
        for (int64_t p = 1; p < intervals[k]; ++p) nv[s0 + p] = v[s0 + p - 1] + d[pointers[s0 + p]];
      }
    }
    v.swap(nv);
  }
  int64_t state = 0;
  for (int64_t s = 1; s < S; ++s)
    if (v[s] > v[state]) state = s;
  *logp_out = v[state];
  for (int64_t t = T - 1; t >= 0; --t) {
    path_out[t] = state;
    const int64_t b = state / per_beat, r = state - b * per_beat;
    int k = 0;  // tempo slot of this state (n_int ~ 42: linear scan is fine)
    while (k + 1 < n_int && first[k + 1] <= r) ++k;
    if (r == first[k]) {
      const int f = back[(static_cast<size_t>(t) * beats + b) * n_int + k];
      state = ((b + beats - 1) % beats) * per_beat + last[f];
    } else {
      state -= 1;
    }
  }
  return BT_OK;
}
