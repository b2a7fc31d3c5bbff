// K12: fused quadratic-phase diagonal pass.  ANY number of commuting 1- and 2-qubit diagonal gates (rz, phaseshift, cz,
// cu1, crz, rzz, the parity diagonals that cnot-rz-cnot sandwiches fold into, ...) is ONE trip through HBM:
//     psi[x] *= exp(2 pi i * (c0 + sum_b alpha_b x_b + sum_{a<b} beta_ab x_a x_b))      x_b = bit b of the amplitude index
// Replaces one apply_unitary_bmm per gate (torchquantum/functional/gate_wrapper.py:96-140) for every gate folded into the
// form; the twiddle stages between the dense 6-qubit blocks of a QFT (Cooley-Tukey: 150 controlled phases per stage
// at n = 31) are the case it was written for.
//
// Phases are uint32 fractions of a turn: sums wrap mod 1 exactly, so hundreds of terms cost no accuracy (the float angle
// handed to sincospif has 24 bits of a turn).  Per CTA: T_ll[4096] = the form restricted to the 12 low index bits (built
// once); per tile of 4096 amplitudes: Q_hh = the form on the tile's high bits (incl. the rank bits of a sharded state),
// v_l = sum_h beta_lh x_h for the 12 low bits, expanded into two 64-entry tables U1 / U2 (bits 0-5 / 6-11); per amplitude:
// turns = Q_hh + T_ll[x_l] + U1[x_l & 63] + U2[x_l >> 6], one sincospif, one complex multiply.  HBM-bound: 16 B per amplitude.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tqb_program.h"

namespace tqb_qp {

constexpr int kBits = TQB_QP_BITS;          // index bits the form may address (local + rank bits)
constexpr int kTileLog = 12;
constexpr int kThreads = 256;

// coef: [0] c0, [1 + b] alpha_b, [1 + kBits + a * kBits + b] beta_ab for a < b (other entries ignored)
__global__ void __launch_bounds__(kThreads, 4)
tqb_quadphase_kernel(float4* __restrict__ state, const uint32_t* __restrict__ coef, int n_local, int n_bits, unsigned rank,
                     long long n_tiles) {
  __shared__ uint32_t s_beta[kBits * kBits];
  __shared__ uint32_t s_alpha[kBits];
  __shared__ uint32_t s_tll[1 << kTileLog];
  __shared__ uint32_t s_r[kBits];
  __shared__ uint32_t s_u1[64], s_u2[64];
  __shared__ uint32_t s_qhh;
  const int tid = threadIdx.x;
  for (int i = tid; i < kBits * kBits; i += kThreads) s_beta[i] = coef[1 + kBits + i];
  if (tid < kBits) s_alpha[tid] = coef[1 + tid];
  const uint32_t c0 = coef[0];
  __syncthreads();
  for (int e = tid; e < (1 << kTileLog); e += kThreads) {
    uint32_t t = 0;
    for (int l = 0; l < kTileLog; ++l)
      if ((e >> l) & 1) {
        t += s_alpha[l];
        for (int m = 0; m < l; ++m)
          if ((e >> m) & 1) t += s_beta[m * kBits + l];
      }
    s_tll[e] = t;
  }
  const unsigned long long tiles_per_state = 1ull << (n_local - kTileLog);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // index bits of this tile above the low 12 (the batch index does not enter: shared parameters only)
    const unsigned long long x = (((unsigned long long)tile & (tiles_per_state - 1ull)) << kTileLog) | ((unsigned long long)rank << n_local);
    __syncthreads();   // the tables of the previous tile are no longer read (and T_ll is complete)
    if (tid < n_bits) {
      uint32_t t = 0;
      if (tid < kTileLog) {
        for (int h = kTileLog; h < n_bits; ++h)
          if ((x >> h) & 1ull) t += s_beta[tid * kBits + h];
      } else if ((x >> tid) & 1ull) {
        t = s_alpha[tid];
        for (int j = kTileLog; j < tid; ++j)
          if ((x >> j) & 1ull) t += s_beta[j * kBits + tid];
      }
      s_r[tid] = t;
    }
    __syncthreads();
    if (tid < 128) {
      const int e = tid & 63, off = tid < 64 ? 0 : 6;
      uint32_t t = 0;
#pragma unroll
      for (int l = 0; l < 6; ++l)
        if ((e >> l) & 1) t += s_r[off + l];
      (tid < 64 ? s_u1 : s_u2)[e] = t;
    } else if (tid == 128) {
      uint32_t q = c0;
      for (int i = kTileLog; i < n_bits; ++i) q += s_r[i];
      s_qhh = q;
    }
    __syncthreads();
    const uint32_t qhh = s_qhh;
    float4* p = state + ((unsigned long long)tile << (kTileLog - 1));
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[tid + kThreads * k];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t xl = 2u * (uint32_t)(tid + kThreads * k);
      const uint32_t hi = s_u2[xl >> 6] + qhh;
      const uint32_t t0 = hi + s_tll[xl] + s_u1[xl & 63u], t1 = hi + s_tll[xl + 1u] + s_u1[(xl & 63u) + 1u];
      float s0, c0f, s1, c1f;
      sincospif((float)(int32_t)t0 * 4.656612873077393e-10f, &s0, &c0f);   // (int32) turns * 2^-31 = the angle in units of pi
      sincospif((float)(int32_t)t1 * 4.656612873077393e-10f, &s1, &c1f);
      const float4 a = v[k];
      float4 o;
      o.x = a.x * c0f - a.y * s0; o.y = a.x * s0 + a.y * c0f;
      o.z = a.z * c1f - a.w * s1; o.w = a.z * s1 + a.w * c1f;
      p[tid + kThreads * k] = o;
    }
  }
}

inline const char* launch_quadphase(float* state, int n_local, long long bsz, const uint32_t* coef, int n_bits, unsigned rank,
                                    int num_sms, cudaStream_t st) {
  if (n_local < kTileLog) return "quadphase needs n_local >= 12";
  if (n_bits < n_local || n_bits > kBits) return "quadphase: n_bits must be in [n_local, TQB_QP_BITS]";
  if (bsz <= 0) return nullptr;
  const long long n_tiles = bsz << (n_local - kTileLog);
  const long long want = (long long)num_sms * 4;
  const unsigned grid = (unsigned)(n_tiles < want ? n_tiles : want);
  tqb_quadphase_kernel<<<grid, kThreads, 0, st>>>(reinterpret_cast<float4*>(state), coef, n_local, n_bits, rank, n_tiles);
  if (cudaGetLastError() != cudaSuccess) return "quadphase launch failed";
  return nullptr;
}

}  // namespace tqb_qp
