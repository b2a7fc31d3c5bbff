// Single-source body of the fused gate-pass kernels.
//
// The same code is compiled twice:
//   * by nvcc for sm_100a (tqb_cuda.cu): Exec = CudaExec, one CTA per tile (or per S samples),
//     `par(f)` runs f(threadIdx.x) followed by __syncthreads();
//   * by g++ for the CPU tests of the host planner (tests/emu/tqb_emu.cpp): Exec = EmuExec,
//     `par(f)` runs f(tid) for every tid of the block in turn.  That build is test
//     infrastructure only; the product never loads it.
// Per-thread state therefore never lives across two `par` phases: amplitudes go back to the
// staged tile at the end of every group.
#pragma once
#include "tqb_program.h"

#if defined(__CUDACC__)
#define TQB_FN __device__ __forceinline__
#else
#include <cmath>
#include <cstring>
#define TQB_FN inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
#endif

namespace tqb {

typedef unsigned int u32;
typedef unsigned long long u64;

TQB_FN float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// conj(a) * b
TQB_FN float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
TQB_FN float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TQB_FN float2 ld2(const float* m) { return make_float2(m[0], m[1]); }

// swizzled position of tile-local amplitude idx in the staged tile (8-byte slots):
// xor the 4 low bits with the next 4 so that the 16 amplitudes a half-warp touches in one
// 64-bit shared access fall in 16 distinct bank pairs for every run of consecutive register bits.
TQB_FN u32 swz(u32 idx) { return idx ^ ((idx >> 4) & 15u); }

struct BlockCtx {
  int T, R, L, n;          // tile bits, register bits, contiguous low bits, local qubits
  int S, tps, nthreads;    // samples per CTA, threads per sample, S * tps
  int bsz, sample0;        // batch size, first sample handled by this CTA
  int multi_tile;          // more than one tile per sample (gradients need atomics)
  u32 tile_ext;            // compressed outer index (+ rank bits in sharded mode)
  u64 outer_off;           // amplitude offset of the tile inside the sample
  u64 sample_stride;       // 2^n
  const int32_t* prog;     // global memory
  const float2* psi_src;   // global state read by the pass (may alias psi)
  const float2* lam_src;   // backward only: dL/dpsi read by the pass
  float2* psi;             // global state written by the pass
  float2* lam;             // backward only: lambda written by the pass
  const float* mats;       // [bsz][mat_stride]
  float* gmats;            // [bsz][mat_stride]
  long long mat_stride;
  // staged (shared) memory
  float2* s_psi; float2* s_lam; float* s_mat; float* s_grad; int32_t* s_prog; u64* s_scat;
};

// ---------------------------------------------------------------------------------------
// tile <-> global
// ---------------------------------------------------------------------------------------
TQB_FN void build_scatter(const BlockCtx& c, int tid) {
  const int nhi = 1 << (c.T - c.L);
  for (int h = tid; h < nhi; h += c.nthreads) {
    u64 off = 0;
    for (int i = 0; i < c.T - c.L; ++i)
      if ((h >> i) & 1) off |= 1ull << c.prog[TQB_H_TBIT + c.L + i];
    c.s_scat[h] = off;
  }
}

TQB_FN void load_tile(const BlockCtx& c, int tid, const float2* g, float2* s) {
  const int slot = tid / c.tps, t = tid % c.tps;
  const int b = c.sample0 + slot;
  const u32 tile = 1u << c.T;
  float2* dst = s + (size_t)slot * tile;
  const u32 lowmask = (1u << c.L) - 1u;
  const float2* src = g + (u64)b * c.sample_stride + c.outer_off;
  for (u32 idx = 2u * t; idx < tile; idx += 2u * c.tps) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < c.bsz) v = *reinterpret_cast<const float4*>(src + ((idx & lowmask) + c.s_scat[idx >> c.L]));
    dst[swz(idx)] = make_float2(v.x, v.y);
    dst[swz(idx + 1)] = make_float2(v.z, v.w);
  }
}

TQB_FN void store_tile(const BlockCtx& c, int tid, float2* g, const float2* s) {
  const int slot = tid / c.tps, t = tid % c.tps;
  const int b = c.sample0 + slot;
  if (b >= c.bsz) return;
  const u32 tile = 1u << c.T;
  const float2* src = s + (size_t)slot * tile;
  const u32 lowmask = (1u << c.L) - 1u;
  float2* dst = g + (u64)b * c.sample_stride + c.outer_off;
  for (u32 idx = 2u * t; idx < tile; idx += 2u * c.tps) {
    float2 a = src[swz(idx)], bq = src[swz(idx + 1)];
    *reinterpret_cast<float4*>(dst + ((idx & lowmask) + c.s_scat[idx >> c.L])) = make_float4(a.x, a.y, bq.x, bq.y);
  }
}

// ---------------------------------------------------------------------------------------
// segment staging: op words, matrices (per sample), zeroed gradient accumulators
// ---------------------------------------------------------------------------------------
template <bool BWD>
TQB_FN void stage_segment(const BlockCtx& c, int tid, int soff) {
  const int ng = c.prog[soff + 0], nops = c.prog[soff + 1];
  const int mat_base = c.prog[soff + 2], mat_len = c.prog[soff + 3];
  const int nwords = TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS + nops * TQB_OP_WORDS;
  for (int i = tid; i < nwords; i += c.nthreads) c.s_prog[i] = c.prog[soff + i];
  for (int i = tid; i < c.S * mat_len; i += c.nthreads) {
    const int slot = i / mat_len, k = i - slot * mat_len;
    const int b = c.sample0 + slot;
    c.s_mat[slot * TQB_SEG_MAX_FLOATS + k] = (b < c.bsz) ? c.mats[(long long)b * c.mat_stride + mat_base + k] : 0.f;
    if (BWD) c.s_grad[slot * TQB_SEG_MAX_FLOATS + k] = 0.f;
  }
}

template <class Exec>
TQB_FN void flush_grads(Exec& ex, const BlockCtx& c, int tid) {
  const int mat_base = c.s_prog[2], mat_len = c.s_prog[3];
  for (int i = tid; i < c.S * mat_len; i += c.nthreads) {
    const int slot = i / mat_len, k = i - slot * mat_len;
    const int b = c.sample0 + slot;
    if (b >= c.bsz) continue;
    const float v = c.s_grad[slot * TQB_SEG_MAX_FLOATS + k];
    float* dst = c.gmats + (long long)b * c.mat_stride + mat_base + k;
    if (c.multi_tile) ex.global_add(dst, v); else *dst = v;
  }
}

// ---------------------------------------------------------------------------------------
// register-resident op application
// ---------------------------------------------------------------------------------------
template <int R>
TQB_FN u32 dep(int e, const u32 (&pm)[R]) {
  u32 d = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) if (e & (1 << j)) d |= pm[j];
  return d;
}

template <int R, int J>
TQB_FN void u1q_fwd(float2 (&x)[1 << R], const float* m, u32 cm, u32 cv, u32 base, const u32 (&pm)[R]) {
  const float2 m00 = ld2(m), m01 = ld2(m + 2), m10 = ld2(m + 4), m11 = ld2(m + 6);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      const float2 a = x[e], b = x[e | (1 << J)];
      x[e] = cadd(cmul(m00, a), cmul(m01, b));
      x[e | (1 << J)] = cadd(cmul(m10, a), cmul(m11, b));
    }
  }
}

// backward of out = A in on one pair: x holds psi_out (unless NOUNAPPLY: psi_in), y holds
// lambda_out.  Leaves psi_in / lambda_in = A^H lambda_out; accumulates g[2i+j] += lam_out_i conj(psi_in_j).
template <int R, int J>
TQB_FN void u1q_bwd(float2 (&x)[1 << R], float2 (&y)[1 << R], const float* m, u32 cm, u32 cv, u32 base,
                    const u32 (&pm)[R], bool grad, bool unapply, float (&g)[8]) {
  const float2 m00 = ld2(m), m01 = ld2(m + 2), m10 = ld2(m + 4), m11 = ld2(m + 6);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      float2 a = x[e], b = x[e | (1 << J)];
      if (unapply) {
        const float2 pa = cadd(cmulc(m00, a), cmulc(m10, b));
        const float2 pb = cadd(cmulc(m01, a), cmulc(m11, b));
        a = pa; b = pb;
        x[e] = a; x[e | (1 << J)] = b;
      }
      const float2 la = y[e], lb = y[e | (1 << J)];
      if (grad) {
        float2 t;
        t = cmulc(a, la); g[0] += t.x; g[1] += t.y;   // G00 = la conj(a)
        t = cmulc(b, la); g[2] += t.x; g[3] += t.y;   // G01 = la conj(b)
        t = cmulc(a, lb); g[4] += t.x; g[5] += t.y;   // G10
        t = cmulc(b, lb); g[6] += t.x; g[7] += t.y;   // G11
      }
      y[e] = cadd(cmulc(m00, la), cmulc(m10, lb));
      y[e | (1 << J)] = cadd(cmulc(m01, la), cmulc(m11, lb));
    }
  }
}

template <int R>
TQB_FN void diag_fwd(float2 (&x)[1 << R], const float* m, u32 pmask, u32 opar, u32 cm, u32 cv, u32 base,
                     const u32 (&pm)[R]) {
  const float2 d0 = ld2(m), d1 = ld2(m + 2);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      const u32 p = (__popc(idx & pmask) + opar) & 1u;
      x[e] = cmul(p ? d1 : d0, x[e]);
    }
  }
}

template <int R>
TQB_FN void diag_bwd(float2 (&x)[1 << R], float2 (&y)[1 << R], const float* m, u32 pmask, u32 opar, u32 cm,
                     u32 cv, u32 base, const u32 (&pm)[R], bool grad, bool unapply, float (&g)[8]) {
  const float2 d0 = ld2(m), d1 = ld2(m + 2);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      const u32 p = (__popc(idx & pmask) + opar) & 1u;
      const float2 d = p ? d1 : d0;
      if (unapply) x[e] = cmulc(d, x[e]);
      if (grad) {
        const float2 t = cmulc(x[e], y[e]);  // lam_out conj(psi_in)
        if (p) { g[2] += t.x; g[3] += t.y; } else { g[0] += t.x; g[1] += t.y; }
      }
      y[e] = cmulc(d, y[e]);
    }
  }
}

// 4x4 on register bits (1,0): quad index q = 2*bit1 + bit0
template <int R>
TQB_FN void u2q_fwd(float2 (&x)[1 << R], const float* m, u32 cm, u32 cv, u32 base, const u32 (&pm)[R]) {
#pragma unroll
  for (int e = 0; e < (1 << R); e += 4) {
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      float2 v[4], o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = x[e + q];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i] = cadd(o[i], cmul(ld2(m + 2 * (4 * i + j)), v[j]));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) x[e + q] = o[q];
    }
  }
}

template <int R>
TQB_FN void u2q_bwd(float2 (&x)[1 << R], float2 (&y)[1 << R], const float* m, u32 cm, u32 cv, u32 base,
                    const u32 (&pm)[R], bool grad, bool unapply, float (&g)[32]) {
#pragma unroll
  for (int e = 0; e < (1 << R); e += 4) {
    const u32 idx = base | dep<R>(e, pm);
    if ((idx & cm) == cv) {
      float2 v[4], l[4], o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { v[q] = x[e + q]; l[q] = y[e + q]; }
      if (unapply) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[j] = cadd(o[j], cmulc(ld2(m + 2 * (4 * i + j)), v[i]));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = o[q]; x[e + q] = o[q]; }
      }
      if (grad) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = cmulc(v[j], l[i]);
            g[2 * (4 * i + j)] += t.x; g[2 * (4 * i + j) + 1] += t.y;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[j] = cadd(o[j], cmulc(ld2(m + 2 * (4 * i + j)), l[i]));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) y[e + q] = o[q];
    }
  }
}

// insert zero bits at the (ascending) positions sp[0..R) of w
template <int R>
TQB_FN u32 insert_zeros(u32 w, const u32 (&sp)[R]) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const u32 low = w & ((1u << sp[j]) - 1u);
    w = ((w >> sp[j]) << (sp[j] + 1)) | low;
  }
  return w;
}

template <int R>
TQB_FN void sort_positions(u32 (&sp)[R]) {
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b + 1 < R - a; ++b)
      if (sp[b] > sp[b + 1]) { const u32 t = sp[b]; sp[b] = sp[b + 1]; sp[b + 1] = t; }
}

template <int R, bool BWD, class Exec>
TQB_FN void run_group(Exec& ex, const BlockCtx& c, int tid, int g) {
  const int ng = c.s_prog[0];
  const int32_t* gh = c.s_prog + TQB_SEG_HDR_WORDS + g * TQB_GROUP_WORDS;
  const int slot = tid / c.tps, t = tid % c.tps;
  u32 pos[R], sp[R], pm[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { pos[j] = (u32)gh[j]; sp[j] = pos[j]; pm[j] = 1u << pos[j]; }
  sort_positions<R>(sp);
  const int nops = gh[4];
  const int32_t* ops = c.s_prog + TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS + gh[5] * TQB_OP_WORDS;
  const int items = 1 << (c.T - R);
  float2* sps = c.s_psi + ((size_t)slot << c.T);
  float2* sla = BWD ? c.s_lam + ((size_t)slot << c.T) : nullptr;
  const float* smat = c.s_mat + slot * TQB_SEG_MAX_FLOATS;
  float* sgrad = BWD ? c.s_grad + slot * TQB_SEG_MAX_FLOATS : nullptr;
  const int width = c.tps < 32 ? c.tps : 32;

  for (int item = t; item < items; item += c.tps) {
    const u32 base = insert_zeros<R>((u32)item, sp);
    float2 x[1 << R];
    float2 y[1 << R];
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      const u32 a = swz(base | dep<R>(e, pm));
      x[e] = sps[a];
      if constexpr (BWD) y[e] = sla[a];
    }
    for (int k = 0; k < nops; ++k) {
      const int32_t* op = ops + (BWD ? (nops - 1 - k) : k) * TQB_OP_WORDS;
      const u32 w0 = (u32)op[TQB_OP_TYPE];
      const u32 cmo = (u32)op[TQB_OP_CM_O], cvo = (u32)op[TQB_OP_CV_O];
      if ((c.tile_ext & cmo) != cvo) continue;  // control on a non-tile (or global) bit: CTA-uniform
      const u32 type = w0 & 0xffu;
      const u32 cm = (u32)op[TQB_OP_CM_T], cv = (u32)op[TQB_OP_CV_T];
      const float* m = smat + op[TQB_OP_SLOT];
      const bool grad = BWD && (w0 & TQB_F_GRAD);
      const bool unapply = !(w0 & TQB_F_NOUNAPPLY);
      if (type == TQB_T_U1Q) {
        const int j = op[TQB_OP_TGT];
        if constexpr (!BWD) {
          if (j == 0) u1q_fwd<R, 0>(x, m, cm, cv, base, pm);
          if constexpr (R > 1) { if (j == 1) u1q_fwd<R, 1>(x, m, cm, cv, base, pm); }
          if constexpr (R > 2) { if (j == 2) u1q_fwd<R, 2>(x, m, cm, cv, base, pm); }
          if constexpr (R > 3) { if (j == 3) u1q_fwd<R, 3>(x, m, cm, cv, base, pm); }
        } else {
          float gacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (j == 0) u1q_bwd<R, 0>(x, y, m, cm, cv, base, pm, grad, unapply, gacc);
          if constexpr (R > 1) { if (j == 1) u1q_bwd<R, 1>(x, y, m, cm, cv, base, pm, grad, unapply, gacc); }
          if constexpr (R > 2) { if (j == 2) u1q_bwd<R, 2>(x, y, m, cm, cv, base, pm, grad, unapply, gacc); }
          if constexpr (R > 3) { if (j == 3) u1q_bwd<R, 3>(x, y, m, cm, cv, base, pm, grad, unapply, gacc); }
          if (grad) ex.template reduce_add<8>(sgrad + op[TQB_OP_SLOT], gacc, width);
        }
      } else if (type == TQB_T_DIAG) {
        const u32 opar = (u32)__popc(c.tile_ext & (u32)op[TQB_OP_PM_O]) & 1u;
        if constexpr (!BWD) {
          diag_fwd<R>(x, m, (u32)op[TQB_OP_PM_T], opar, cm, cv, base, pm);
        } else {
          float gacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          diag_bwd<R>(x, y, m, (u32)op[TQB_OP_PM_T], opar, cm, cv, base, pm, grad, unapply, gacc);
          if (grad) ex.template reduce_add<4>(sgrad + op[TQB_OP_SLOT], gacc, width);
        }
      } else {  // TQB_T_U2Q
        if constexpr (R >= 2) {
          if constexpr (!BWD) {
            u2q_fwd<R>(x, m, cm, cv, base, pm);
          } else {
            float g32[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) g32[i] = 0.f;
            u2q_bwd<R>(x, y, m, cm, cv, base, pm, grad, unapply, g32);
            if (grad) ex.template reduce_add<32>(sgrad + op[TQB_OP_SLOT], g32, width);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      const u32 a = swz(base | dep<R>(e, pm));
      sps[a] = x[e];
      if constexpr (BWD) sla[a] = y[e];
    }
  }
}

// whole block: load tile(s), run every segment/group, store tile(s)
template <int R, bool BWD, class Exec>
TQB_FN void run_block(Exec& ex, const BlockCtx& c) {
  ex.par([&](int tid) { build_scatter(c, tid); });
  ex.par([&](int tid) {
    load_tile(c, tid, c.psi_src, c.s_psi);
    if (BWD) load_tile(c, tid, c.lam_src, c.s_lam);
  });
  const int nseg = c.prog[TQB_H_NSEG];
  for (int si = 0; si < nseg; ++si) {
    const int s = BWD ? nseg - 1 - si : si;
    const int soff = c.prog[TQB_H_SEGOFF + s];
    ex.par([&](int tid) { stage_segment<BWD>(c, tid, soff); });
    const int ng = c.s_prog[0];
    for (int gi = 0; gi < ng; ++gi) {
      const int g = BWD ? ng - 1 - gi : gi;
      ex.par([&](int tid) { run_group<R, BWD>(ex, c, tid, g); });
    }
    if (BWD) ex.par([&](int tid) { flush_grads(ex, c, tid); });
  }
  ex.par([&](int tid) {
    store_tile(c, tid, c.psi, c.s_psi);
    if (BWD) store_tile(c, tid, c.lam, c.s_lam);
  });
}

// offset (in amplitudes) of tile `tile_id` inside a sample: deposit its bits into the non-tile positions
TQB_FN u64 outer_offset(const int32_t* prog, int n, u32 tile_id) {
  const int T = prog[TQB_H_T];
  u64 tmask = 0;
  for (int i = 0; i < T; ++i) tmask |= 1ull << prog[TQB_H_TBIT + i];
  u64 off = 0;
  int k = 0;
  for (int p = 0; p < n; ++p)
    if (!((tmask >> p) & 1ull)) { if ((tile_id >> k) & 1u) off |= 1ull << p; ++k; }
  return off;
}

// shared-memory carve-up (bytes) for a block of S samples of 2^T amplitudes
struct SmemLayout { size_t psi, lam, mat, grad, prog, scat, total; };
inline
#if defined(__CUDACC__)
__host__ __device__
#endif
SmemLayout smem_layout(int T, int L, int S, bool bwd) {
  SmemLayout l;
  size_t off = 0;
  const size_t tile = ((size_t)S << T) * sizeof(float2);
  l.psi = off; off += tile;
  l.lam = off; if (bwd) off += tile;
  l.mat = off; off += (size_t)S * TQB_SEG_MAX_FLOATS * sizeof(float);
  l.grad = off; if (bwd) off += (size_t)S * TQB_SEG_MAX_FLOATS * sizeof(float);
  l.prog = off; off += (size_t)TQB_SEG_MAX_WORDS * sizeof(int32_t);
  off = (off + 15) & ~(size_t)15;
  l.scat = off; off += ((size_t)1 << (T - L)) * sizeof(u64);
  l.total = (off + 15) & ~(size_t)15;
  return l;
}

// ---------------------------------------------------------------------------------------
// parameter -> matrix builders, written once over a scalar type (float, or a dual number for
// the analytic derivative used by the backward builder)
// ---------------------------------------------------------------------------------------
struct Dual { float v, d; };
TQB_FN Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, a.d + b.d}; }
TQB_FN Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, a.d - b.d}; }
TQB_FN Dual operator*(Dual a, Dual b) { return Dual{a.v * b.v, a.v * b.d + a.d * b.v}; }
TQB_FN Dual operator-(Dual a) { return Dual{-a.v, -a.d}; }
TQB_FN Dual s_const(Dual, float c) { return Dual{c, 0.f}; }
TQB_FN float s_const(float, float c) { return c; }
TQB_FN void s_sincos(float a, float& s, float& c) {
#if defined(__CUDA_ARCH__)
  sincosf(a, &s, &c);
#else
  s = sinf(a); c = cosf(a);
#endif
}
TQB_FN void s_sincos(Dual a, Dual& s, Dual& c) {
  float sv, cv;
  s_sincos(a.v, sv, cv);
  s = Dual{sv, cv * a.d};
  c = Dual{cv, -sv * a.d};
}

template <class T> struct Cx { T re, im; };
template <class T> TQB_FN Cx<T> operator*(Cx<T> a, Cx<T> b) { return Cx<T>{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <class T> TQB_FN Cx<T> scale(Cx<T> a, T s) { return Cx<T>{a.re * s, a.im * s}; }
template <class T> TQB_FN Cx<T> neg(Cx<T> a) { return Cx<T>{-a.re, -a.im}; }
template <class T> TQB_FN Cx<T> expi(T a) { T s, c; s_sincos(a, s, c); return Cx<T>{c, s}; }
template <class T> TQB_FN Cx<T> mul_mi(Cx<T> a) { return Cx<T>{a.im, -a.re}; }  // -i * a

// builds the *un-inverted* matrix U(theta) row-major into u (4, 2 or 16 entries); returns #entries
template <class T>
TQB_FN int build_gate(int kind, const T* th, Cx<T>* u) {
  const T zero = s_const(th[0], 0.f), one = s_const(th[0], 1.f), half = s_const(th[0], 0.5f);
  const Cx<T> cz = Cx<T>{zero, zero}, c1 = Cx<T>{one, zero};
  T s, c;
  switch (kind) {
    case TQB_K_RX: {
      s_sincos(th[0] * half, s, c);
      u[0] = Cx<T>{c, zero}; u[1] = Cx<T>{zero, -s}; u[2] = Cx<T>{zero, -s}; u[3] = Cx<T>{c, zero};
      return 4;
    }
    case TQB_K_RY: {
      s_sincos(th[0] * half, s, c);
      u[0] = Cx<T>{c, zero}; u[1] = Cx<T>{-s, zero}; u[2] = Cx<T>{s, zero}; u[3] = Cx<T>{c, zero};
      return 4;
    }
    case TQB_K_ROT: {  // th = (phi, theta, omega)
      s_sincos(th[1] * half, s, c);
      const T a = (th[0] + th[2]) * half, b = (th[0] - th[2]) * half;
      u[0] = scale(expi(-a), c); u[1] = neg(scale(expi(b), s));
      u[2] = scale(expi(-b), s); u[3] = scale(expi(a), c);
      return 4;
    }
    case TQB_K_U2: case TQB_K_U2RAW: {  // th = (phi, lam)
      const T k = s_const(th[0], kind == TQB_K_U2 ? 0.70710678118654752440f : 1.0f);
      u[0] = Cx<T>{k, zero}; u[1] = neg(scale(expi(th[1]), k));
      u[2] = scale(expi(th[0]), k); u[3] = scale(expi(th[0] + th[1]), k);
      return 4;
    }
    case TQB_K_U3: {  // th = (theta, phi, lam)
      s_sincos(th[0] * half, s, c);
      u[0] = Cx<T>{c, zero}; u[1] = neg(scale(expi(th[2]), s));
      u[2] = scale(expi(th[1]), s); u[3] = scale(expi(th[1] + th[2]), c);
      return 4;
    }
    case TQB_K_CU: {  // th = (theta, phi, lam, gamma)
      s_sincos(th[0] * half, s, c);
      u[0] = scale(expi(th[3]), c); u[1] = neg(scale(expi(th[2] + th[3]), s));
      u[2] = scale(expi(th[1] + th[3]), s); u[3] = scale(expi(th[1] + th[2] + th[3]), c);
      return 4;
    }
    case TQB_K_R: {  // th = (theta, phi): [[c, -i s e^{-i phi}], [-i s e^{i phi}, c]]
      s_sincos(th[0] * half, s, c);
      u[0] = Cx<T>{c, zero}; u[1] = mul_mi(scale(expi(-th[1]), s));
      u[2] = mul_mi(scale(expi(th[1]), s)); u[3] = Cx<T>{c, zero};
      return 4;
    }
    case TQB_K_RZ: {
      u[0] = expi(-(th[0] * half)); u[1] = expi(th[0] * half);
      return 2;
    }
    case TQB_K_PHASE: { u[0] = c1; u[1] = expi(th[0]); return 2; }
    case TQB_K_GPHASE: { u[0] = expi(th[0]); u[1] = u[0]; return 2; }
    default: break;
  }
  for (int i = 0; i < 16; ++i) u[i] = cz;
  switch (kind) {
    case TQB_K_RXX: case TQB_K_RYY: {
      s_sincos(th[0] * half, s, c);
      const Cx<T> d = Cx<T>{c, zero}, mis = Cx<T>{zero, -s}, pis = Cx<T>{zero, s};
      u[0] = d; u[5] = d; u[10] = d; u[15] = d;
      u[3] = kind == TQB_K_RXX ? mis : pis; u[12] = u[3]; u[6] = mis; u[9] = mis;
      return 16;
    }
    case TQB_K_RZX: {
      s_sincos(th[0] * half, s, c);
      const Cx<T> d = Cx<T>{c, zero}, mis = Cx<T>{zero, -s}, pis = Cx<T>{zero, s};
      u[0] = d; u[1] = mis; u[4] = mis; u[5] = d; u[10] = d; u[11] = pis; u[14] = pis; u[15] = d;
      return 16;
    }
    case TQB_K_SEXC: {
      s_sincos(th[0] * half, s, c);
      u[0] = c1; u[15] = c1; u[5] = Cx<T>{c, zero}; u[6] = Cx<T>{-s, zero}; u[9] = Cx<T>{s, zero}; u[10] = Cx<T>{c, zero};
      return 16;
    }
    case TQB_K_XXMYY: {  // th = (theta, beta)
      s_sincos(th[0] * half, s, c);
      u[0] = Cx<T>{c, zero}; u[15] = Cx<T>{c, zero}; u[5] = c1; u[10] = c1;
      u[3] = mul_mi(scale(expi(-th[1]), s)); u[12] = mul_mi(scale(expi(th[1]), s));
      return 16;
    }
    case TQB_K_XXPYY: {
      s_sincos(th[0] * half, s, c);
      u[0] = c1; u[15] = c1; u[5] = Cx<T>{c, zero}; u[10] = Cx<T>{c, zero};
      u[6] = mul_mi(scale(expi(th[1]), s)); u[9] = mul_mi(scale(expi(-th[1]), s));
      return 16;
    }
    default: break;
  }
  return 0;
}

TQB_FN int kind_dim(int nent) { return nent == 16 ? 4 : (nent == 4 ? 2 : 0); }  // 0: diagonal pair

// one (op, sample): params -> applied matrix (inverse folded in) written to the sample's matrix row
TQB_FN void build_one(const int32_t* rec, int b, const float* params_sh, const float* params_ps, int ps_stride,
                      const float* consts, float* mats_row) {
  const int kind = rec[TQB_B_KIND];
  float* out = mats_row + rec[TQB_B_MOFF];
  if (kind == TQB_K_EXTERN) return;
  if (kind == TQB_K_CONST) {
    for (int i = 0; i < rec[TQB_B_NFLOAT]; ++i) out[i] = consts[rec[TQB_B_COFF] + i];
    return;
  }
  const bool inv = rec[TQB_B_FLAGS] & TQB_BF_INVERSE;
  const float* p = (rec[TQB_B_FLAGS] & TQB_BF_PER_SAMPLE) ? params_ps + (long long)b * ps_stride + rec[TQB_B_POFF]
                                                          : params_sh + rec[TQB_B_POFF];
  float th[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < rec[TQB_B_NPARAM]; ++i) th[i] = p[i];
  Cx<float> u[16];
  const int nent = build_gate<float>(kind, th, u);
  const int d = kind_dim(nent);
  if (d == 0) {
    for (int i = 0; i < 2; ++i) { out[2 * i] = u[i].re; out[2 * i + 1] = inv ? -u[i].im : u[i].im; }
  } else {
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        const Cx<float> v = inv ? u[j * d + i] : u[i * d + j];
        out[2 * (i * d + j)] = v.re;
        out[2 * (i * d + j) + 1] = inv ? -v.im : v.im;
      }
  }
}

// one (op, sample): dL/dtheta_p = Re sum conj(dA_ij/dtheta_p) G_ij with A the applied matrix
TQB_FN void build_one_bwd(const int32_t* rec, int b, const float* params_sh, const float* params_ps, int ps_stride,
                          const float* gmats_row, float* gsh_row, float* gps_row) {
  const int kind = rec[TQB_B_KIND];
  if (kind == TQB_K_EXTERN || kind == TQB_K_CONST) return;
  const bool inv = rec[TQB_B_FLAGS] & TQB_BF_INVERSE;
  const bool ps = rec[TQB_B_FLAGS] & TQB_BF_PER_SAMPLE;
  const float* p = ps ? params_ps + (long long)b * ps_stride + rec[TQB_B_POFF] : params_sh + rec[TQB_B_POFF];
  const float* G = gmats_row + rec[TQB_B_MOFF];
  const int np = rec[TQB_B_NPARAM];
  for (int q = 0; q < np; ++q) {
    Dual th[4];
    for (int i = 0; i < 4; ++i) th[i] = Dual{i < np ? p[i] : 0.f, i == q ? 1.f : 0.f};
    Cx<Dual> u[16];
    const int nent = build_gate<Dual>(kind, th, u);
    const int d = kind_dim(nent);
    float acc = 0.f;
    if (d == 0) {
      // applied a_p = d_p (or conj d_p when inverted): Re conj(da) G = dre*Gre + dim*Gim ; inverse flips dim
      for (int i = 0; i < 2; ++i) acc += u[i].re.d * G[2 * i] + (inv ? -u[i].im.d : u[i].im.d) * G[2 * i + 1];
    } else {
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          // applied A_ij = U_ij, or conj(U_ji) when inverted
          const Cx<Dual> v = inv ? u[j * d + i] : u[i * d + j];
          acc += v.re.d * G[2 * (i * d + j)] + (inv ? -v.im.d : v.im.d) * G[2 * (i * d + j) + 1];
        }
    }
    (ps ? gps_row : gsh_row)[rec[TQB_B_POFF] + q] = acc;
  }
}

}  // namespace tqb
