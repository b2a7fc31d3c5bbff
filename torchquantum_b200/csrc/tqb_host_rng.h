// Host-side Mersenne Twister stream of CPython's `random` module, in C.
//
// The reference samples bitstrings on the host with `random.choices(..., k=n_shots)`
// (torchquantum/measurement/measurements.py:37-74), which draws one `random.random()` per shot from the global MT19937
// state.  Histograms can only be bit-exact with the reference if the same stream is consumed, so the uniforms stay a host
// product; what changes is that they are generated here at ~2 ns each instead of one interpreter call each (a million
// shots: 2 ms instead of 150 ms) and the interpreter's state is advanced by exactly the numbers consumed.
//
// state: 625 words as `random.getstate()[1]` returns them (624 words of MT state + the position).
#pragma once
#include <stdint.h>

static inline void tqb_mt19937_refill(uint32_t* mt) {
  const int N = 624, M = 397;
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  int kk = 0;
  for (; kk < N - M; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; kk < N - 1; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  const uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

static inline uint32_t tqb_mt19937_next(uint32_t* state625) {
  uint32_t pos = state625[624];
  if (pos >= 624u) {
    tqb_mt19937_refill(state625);
    pos = 0;
  }
  uint32_t y = state625[pos];
  state625[624] = pos + 1;
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// out[i] = float32(random.random()) for i < n: 53-bit doubles from two words each, as _randommodule.c builds them
static inline void tqb_mt19937_uniforms(uint32_t* state625, long long n, float* out) {
  for (long long i = 0; i < n; ++i) {
    const uint32_t a = tqb_mt19937_next(state625) >> 5, b = tqb_mt19937_next(state625) >> 6;
    out[i] = (float)(((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0));
  }
}
