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
#include <type_traits>

#include "tqb_program.h"

// compile-time op-type mask for experiments ("lean" builds holding only the op types one program family needs ran
// 2-3 % faster: not worth extra binaries; profiles/r01_v3_notes.txt).  Default: every type.
#ifndef TQB_TYPES
#define TQB_TYPES 0xffffu
#endif
#define TQB_HAS(T) (((TQB_TYPES) >> (T)) & 1u)
#if defined(__CUDACC__)
#define TQB_FN __device__ __forceinline__
#define TQB_COLD __device__ __noinline__
#define TQB_FMA(a, b, c) fmaf((a), (b), (c))
#else
#include <cmath>
#include <cstring>
#define TQB_FN inline
#define TQB_COLD static
#define TQB_FMA(a, b, c) ((a) * (b) + (c))
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
#endif

namespace tqb {

typedef unsigned int u32;
typedef unsigned long long u64;

TQB_FN float2 cmul(float2 a, float2 b) { return make_float2(TQB_FMA(-a.y, b.y, a.x * b.x), TQB_FMA(a.y, b.x, a.x * b.y)); }
// conj(a) * b
TQB_FN float2 cmulc(float2 a, float2 b) { return make_float2(TQB_FMA(a.y, b.y, a.x * b.x), TQB_FMA(-a.y, b.x, a.x * b.y)); }
TQB_FN float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TQB_FN float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// m0 * a + m1 * b  (8 FMA-pipe instructions)
TQB_FN float2 cfma2(float2 m0, float2 a, float2 m1, float2 b) {
  float re = m0.x * a.x, im = m0.x * a.y;
  re = TQB_FMA(-m0.y, a.y, re); im = TQB_FMA(m0.y, a.x, im);
  re = TQB_FMA(m1.x, b.x, re);  im = TQB_FMA(m1.x, b.y, im);
  re = TQB_FMA(-m1.y, b.y, re); im = TQB_FMA(m1.y, b.x, im);
  return make_float2(re, im);
}
TQB_FN float4 ld4(const float* m) { return *reinterpret_cast<const float4*>(m); }
// an op's staged matrix: pointer + its first four floats, fetched one op ahead of use (run_group)
struct Mat { const float* p; float4 q0; };

// swizzled position of tile-local amplitude idx in the staged tile (8-byte slots); two layouts, chosen per launch:
// pair = true  (states streamed through HBM): xor index bits 4-6 into bits 1-3.  Bit 0 stays, so the amplitude pair
//   (2p, 2p+1) is one aligned 16-byte slot pair: tiles are filled with 16-byte cp.async and drained with 128-bit
//   shared loads (half the copy instructions: a 1-gate pass went from 4.9 to 6.1 TB/s); the 64-bit accesses of a
//   half-warp fall in >= 8 distinct bank pairs for runs of consecutive register bits.
// pair = false (a whole sample staged per CTA, hundreds of gates per load): xor bits 4-7 into bits 0-3, 16 distinct
//   bank pairs (the on-chip config-2 passes are 2 % faster with it; their tile IO does not matter).
// (GF(2)-linear either way: swz(a ^ b) == swz(a) ^ swz(b).)
TQB_FN u32 swz(u32 idx, bool pair) { return pair ? idx ^ (((idx >> 4) & 7u) << 1) : idx ^ ((idx >> 4) & 15u); }

struct BlockCtx {
  int T, L;                // tile bits, contiguous low bits
  int S, tps, nthreads;    // samples per CTA, threads per sample, S * tps
  int bsz, sample0;        // batch size, first sample handled by this work item
  int multi_tile;          // more than one tile per sample (gradients need atomics; pair swizzle)
  u32 tile_ext;            // compressed outer index (+ rank bits in sharded mode)
  u64 outer_off;           // amplitude offset of the tile inside the sample
  u64 sample_stride;       // 2^n
  const int32_t* prog;     // global memory
  const float2* psi_src;   // global state read by the pass (may alias psi)
  const float2* lam_src;   // backward only: dL/dpsi read by the pass
  float2* psi;             // global state written by the pass
  float2* lam;             // backward only: lambda written by the pass
  const float* msh;        // shared matrix row
  const float* mps;        // [bsz][mps_stride] per-sample matrix rows
  float* gr;               // [bsz][gr_stride] gradient rows
  long long mps_stride, gr_stride;
  int ps_cap, g_cap;       // staged per-sample floats (stride of s_mps / s_grad per sample slot)
  // staged (shared) memory
  float2* s_psi; float2* s_lam; float* s_msh; float* s_mps; float* s_grad; int32_t* s_prog; u64* s_scat;
  float* s_part;           // backward: per-thread gradient partials [g_cap][nthreads + 1]
  int scale_sh, scale_ps;  // float offsets of the pass's deferred scale factor in msh / mps rows (-1: none)
  u64* s_olut;             // [3][256]: amplitude offset contributed by byte j of the tile id (streamed states only)
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

// per-CTA lookup tables: s_scat (tile-local high bits -> amplitude offset) and, for streamed states,
// s_olut (byte j of the tile id -> amplitude offset of the non-tile bits it selects)
TQB_FN void build_tables(const BlockCtx& c, int tid, int n, bool streamed) {
  build_scatter(c, tid);
  if (!streamed) return;
  u64 tmask = 0;
  for (int i = 0; i < c.T; ++i) tmask |= 1ull << c.prog[TQB_H_TBIT + i];
  for (int e = tid; e < 3 * 256; e += c.nthreads) {
    const int byte = e >> 8, v = e & 255;
    u64 off = 0;
    int k = 0;
    for (int p = 0; p < n; ++p)
      if (!((tmask >> p) & 1ull)) {
        if ((k >> 3) == byte && ((v >> (k & 7)) & 1)) off |= 1ull << p;
        ++k;
      }
    c.s_olut[e] = off;
  }
}

// asynchronous (cp.async on the GPU) copy of the work item's tile(s) into the staged buffer
template <class Exec>
TQB_FN void load_tile_async(Exec& ex, const BlockCtx& c, int tid, const float2* g, float2* s) {
  const int slot = tid / c.tps, t = tid % c.tps;
  const int b = c.sample0 + slot;
  const u32 tile = 1u << c.T;
  float2* dst = s + (size_t)slot * tile;
  const u32 lowmask = (1u << c.L) - 1u;
  if (b >= c.bsz) {
    for (u32 idx = t; idx < tile; idx += c.tps) dst[idx] = make_float2(0.f, 0.f);
    return;
  }
  const float2* src = g + (u64)b * c.sample_stride + c.outer_off;
  const bool pair = c.multi_tile;
  if (pair && c.L >= 1) {
    if (c.tps == 256 && tile >= 512u) {
      // pair p = t + 256 k covers idx = 2t + 512 k (+1): the swizzle only mixes bits 1-6, so swz(idx) = swz(2t) + 512 k
      const float2* src_t = src + ((2u * (u32)t) & lowmask);
      float2* dst_t = dst + swz(2u * (u32)t, true);
#pragma unroll 4
      for (u32 k = 0; k < tile; k += 512u)
        ex.async_copy16(dst_t + k, src_t + ((k & lowmask) + c.s_scat[(2u * (u32)t + k) >> c.L]));
      return;
    }
    for (u32 idx = 2u * t; idx < tile; idx += 2u * c.tps)
      ex.async_copy16(dst + swz(idx, true), src + ((idx & lowmask) + c.s_scat[idx >> c.L]));
    return;
  }
  if (!pair && c.tps == 256) {
    // idx = t + 256 k: the swizzle only mixes bits 0-7, so swz(idx) = swz(t) + 256 k
    const float2* src_t = src + ((u32)t & lowmask);
    float2* dst_t = dst + swz((u32)t, false);
#pragma unroll 4
    for (u32 k = 0; k < tile; k += 256u) ex.async_copy8(dst_t + k, src_t + ((k & lowmask) + c.s_scat[((u32)t + k) >> c.L]));
    return;
  }
  for (u32 idx = t; idx < tile; idx += c.tps)
    ex.async_copy8(dst + swz(idx, pair), src + ((idx & lowmask) + c.s_scat[idx >> c.L]));
}

// deferred scale factor of the pass for sample b (scaled op forms, csrc/tqb_program.h)
TQB_FN bool pass_scale(const BlockCtx& c, int b, float2& f) {
  if (c.scale_sh < 0 && c.scale_ps < 0) return false;
  f = make_float2(1.f, 0.f);
  if (c.scale_sh >= 0) f = make_float2(c.msh[c.scale_sh], c.msh[c.scale_sh + 1]);
  if (c.scale_ps >= 0) {
    const float* p = c.mps + (long long)b * c.mps_stride + c.scale_ps;
    f = cmul(f, make_float2(p[0], p[1]));
  }
  return true;
}

TQB_FN void store_tile(const BlockCtx& c, int tid, float2* g, const float2* s) {
  const int slot = tid / c.tps, t = tid % c.tps;
  const int b = c.sample0 + slot;
  if (b >= c.bsz) return;
  const u32 tile = 1u << c.T;
  const float2* src = s + (size_t)slot * tile;
  const u32 lowmask = (1u << c.L) - 1u;
  float2* dst = g + (u64)b * c.sample_stride + c.outer_off;
  float2 f;
  const bool scaled = pass_scale(c, b, f);
  if (c.tps == 256 && c.L >= 1) {
    // pair p = t + 256 k covers idx = 2p, 2p + 1 = 2t + 512 k (+1): swz(idx) = swz(2t) + 512 k
    const bool pair = c.multi_tile;
    const u32 s0 = swz(2u * (u32)t, pair);
    const float2* src_a = src + s0;
    const float2* src_b = src + (s0 ^ 1u);
    float2* dst_t = dst + ((2u * (u32)t) & lowmask);
    if (pair && !scaled) {
#pragma unroll 4
      for (u32 k = 0; k < tile; k += 512u)
        *reinterpret_cast<float4*>(dst_t + ((k & lowmask) + c.s_scat[(2u * (u32)t + k) >> c.L])) =
            *reinterpret_cast<const float4*>(src_a + k);
      return;
    }
    if (scaled) {
#pragma unroll 4
      for (u32 k = 0; k < tile; k += 512u) {
        const float2 a = cmul(f, src_a[k]), bq = cmul(f, src_b[k]);
        *reinterpret_cast<float4*>(dst_t + ((k & lowmask) + c.s_scat[(2u * (u32)t + k) >> c.L])) = make_float4(a.x, a.y, bq.x, bq.y);
      }
      return;
    }
#pragma unroll 4
    for (u32 k = 0; k < tile; k += 512u) {
      const float2 a = src_a[k], bq = src_b[k];
      *reinterpret_cast<float4*>(dst_t + ((k & lowmask) + c.s_scat[(2u * (u32)t + k) >> c.L])) = make_float4(a.x, a.y, bq.x, bq.y);
    }
    return;
  }
  if (c.L >= 1) {
    for (u32 idx = 2u * t; idx < tile; idx += 2u * c.tps) {
      float2 a = src[swz(idx, c.multi_tile)], bq = src[swz(idx + 1, c.multi_tile)];
      if (scaled) { a = cmul(f, a); bq = cmul(f, bq); }
      *reinterpret_cast<float4*>(dst + ((idx & lowmask) + c.s_scat[idx >> c.L])) = make_float4(a.x, a.y, bq.x, bq.y);
    }
  } else {
    for (u32 idx = t; idx < tile; idx += c.tps) {
      float2 a = src[swz(idx, c.multi_tile)];
      if (scaled) a = cmul(f, a);
      dst[c.s_scat[idx]] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------
// segment staging: op words + shared matrices (once per CTA and segment), per-sample matrices
// and zeroed gradient accumulators (once per work item and segment)
// ---------------------------------------------------------------------------------------
TQB_FN void stage_segment_shared(const BlockCtx& c, int tid, int soff) {
  const int ng = c.prog[soff + TQB_S_NGROUPS], nops = c.prog[soff + TQB_S_NOPS];
  const int sh_base = c.prog[soff + TQB_S_SH_BASE], sh_len = c.prog[soff + TQB_S_SH_LEN];
  const int nwords = TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS + nops * TQB_OP_WORDS;
  for (int i = tid; i < nwords; i += c.nthreads) c.s_prog[i] = c.prog[soff + i];
  for (int i = tid; i < sh_len; i += c.nthreads) c.s_msh[i] = c.msh[sh_base + i];
}

template <bool BWD>
TQB_FN void stage_segment_sample(const BlockCtx& c, int tid) {
  const int ps_base = c.s_prog[TQB_S_PS_BASE], ps_len = c.s_prog[TQB_S_PS_LEN];
  for (int i = tid; i < c.S * ps_len; i += c.nthreads) {
    const int slot = i / ps_len, k = i - slot * ps_len;
    const int b = c.sample0 + slot;
    c.s_mps[slot * c.ps_cap + k] = (b < c.bsz) ? c.mps[(long long)b * c.mps_stride + ps_base + k] : 0.f;
  }
  if (BWD) {
    // every thread owns column `tid` of the partial-gradient matrix: zero it without any synchronisation
    const int g_len = c.s_prog[TQB_S_G_LEN];
    float* col = c.s_part + tid;
    for (int k = 0; k < g_len; ++k) col[k * (c.nthreads + 1)] = 0.f;
  }
}

// The backward sweep accumulates every op's gradient partials per THREAD (stash) and sums them over the threads
// of a sample once per segment: thread (slot, k) adds the tps partials of float k in thread order.  (One warp-
// shuffle reduction + shared atomics per op and item cost more than the op's arithmetic on the small batched
// states: 5 dependent shuffle steps with 3 warps per scheduler to hide them.)  Row stride nthreads + 1: both
// the stash (consecutive tid) and this sum (consecutive k) are bank-conflict free; the order is deterministic.
template <int N>
TQB_FN void stash(float* col, int pstride, int gslot, const float* v) {
#pragma unroll
  for (int i = 0; i < N; ++i) col[(gslot + i) * pstride] += v[i];
}

TQB_FN void reduce_partials(const BlockCtx& c, int tid) {
  const int g_len = c.s_prog[TQB_S_G_LEN];
  const int pstride = c.nthreads + 1;
  for (int i = tid; i < c.S * g_len; i += c.nthreads) {
    const int slot = i / g_len, k = i - slot * g_len;
    const float* row = c.s_part + k * pstride + slot * c.tps;
    float acc = 0.f;
    for (int t = 0; t < c.tps; ++t) acc += row[t];
    c.s_grad[slot * c.g_cap + k] = acc;
  }
}

template <class Exec>
TQB_FN void flush_grads(Exec& ex, const BlockCtx& c, int tid) {
  const int g_base = c.s_prog[TQB_S_G_BASE], g_len = c.s_prog[TQB_S_G_LEN];
  for (int i = tid; i < c.S * g_len; i += c.nthreads) {
    const int slot = i / g_len, k = i - slot * g_len;
    const int b = c.sample0 + slot;
    if (b >= c.bsz) continue;
    const float v = c.s_grad[slot * c.g_cap + k];
    float* dst = c.gr + (long long)b * c.gr_stride + g_base + k;
    if (c.multi_tile) ex.global_add(dst, v); else *dst = v;
  }
}

// ---------------------------------------------------------------------------------------
// register-resident op application.  x[e]: the 2^R amplitudes of this thread, e = pattern of the
// register bits.  `act`: bit e set when element e takes part (controls satisfied); only looked at
// when MASKED.  ADJ applies the conjugate transpose (backward sweep).
// ---------------------------------------------------------------------------------------
// The pair updates below are written "in place": every output is produced by the instruction that
// reads the last use of the register it replaces, so the register allocator needs no moves where
// the op cases merge again (with the naive order the switch in apply_op_* costs ~2 MOV per amplitude).
TQB_FN void pair_gen(float2& a, float2& b, float2 m00, float2 m01, float2 m10, float2 m11) {
  // contribution of b to a'
  float tr = m01.x * b.x, ti = m01.x * b.y;
  tr = TQB_FMA(-m01.y, b.y, tr); ti = TQB_FMA(m01.y, b.x, ti);
  // b' = m10 a + m11 b
  const float p = m11.y * b.x;
  float br = m11.x * b.x;
  br = TQB_FMA(-m11.y, b.y, br);
  float bi = TQB_FMA(m11.x, b.y, p);
  br = TQB_FMA(m10.x, a.x, br); bi = TQB_FMA(m10.x, a.y, bi);
  br = TQB_FMA(-m10.y, a.y, br); bi = TQB_FMA(m10.y, a.x, bi);
  // a' = m00 a + t
  const float q = TQB_FMA(m00.y, a.x, ti);
  float ar = TQB_FMA(m00.x, a.x, tr);
  ar = TQB_FMA(-m00.y, a.y, ar);
  const float ai = TQB_FMA(m00.x, a.y, q);
  a = make_float2(ar, ai);
  b = make_float2(br, bi);
}

TQB_FN void pair_real(float2& a, float2& b, float m00, float m01, float m10, float m11) {
  const float tr = m01 * b.x, ti = m01 * b.y;
  float br = m11 * b.x, bi = m11 * b.y;
  br = TQB_FMA(m10, a.x, br); bi = TQB_FMA(m10, a.y, bi);
  a = make_float2(TQB_FMA(m00, a.x, tr), TQB_FMA(m00, a.y, ti));
  b = make_float2(br, bi);
}

// [[c, -is], [-is, c]]
TQB_FN void pair_rx(float2& a, float2& b, float cc, float s) {
  const float t1 = s * b.y, t2 = -s * b.x;
  float br = cc * b.x, bi = cc * b.y;
  br = TQB_FMA(s, a.y, br); bi = TQB_FMA(-s, a.x, bi);
  a = make_float2(TQB_FMA(cc, a.x, t1), TQB_FMA(cc, a.y, t2));
  b = make_float2(br, bi);
}

TQB_FN void cmul_inplace(float2 d, float2& v) {
  const float t = d.y * v.x;
  float r = d.x * v.x;
  r = TQB_FMA(-d.y, v.y, r);
  v = make_float2(r, TQB_FMA(d.x, v.y, t));
}

// ---- scaled forms: one FMA per float; the common factor goes to the pass's deferred scale -------------
// mode 0: [[1, -ir], [-ir, 1]]   mode 1: [[ir, 1], [1, ir]]
template <int MODE>
TQB_FN void pair_rxs(float2& a, float2& b, float r) {
  if (MODE == 0) {
    const float ax = TQB_FMA(r, b.y, a.x), ay = TQB_FMA(-r, b.x, a.y);
    b = make_float2(TQB_FMA(r, a.y, b.x), TQB_FMA(-r, a.x, b.y));
    a = make_float2(ax, ay);
  } else {
    const float ax = TQB_FMA(-r, a.y, b.x), ay = TQB_FMA(r, a.x, b.y);
    b = make_float2(TQB_FMA(-r, b.y, a.x), TQB_FMA(r, b.x, a.y));
    a = make_float2(ax, ay);
  }
}
// mode 0: [[1, -r], [r, 1]]   mode 1: [[r, -1], [1, r]]
template <int MODE>
TQB_FN void pair_rys(float2& a, float2& b, float r) {
  if (MODE == 0) {
    const float ax = TQB_FMA(-r, b.x, a.x), ay = TQB_FMA(-r, b.y, a.y);
    b = make_float2(TQB_FMA(r, a.x, b.x), TQB_FMA(r, a.y, b.y));
    a = make_float2(ax, ay);
  } else {
    const float ax = TQB_FMA(r, a.x, -b.x), ay = TQB_FMA(r, a.y, -b.y);
    b = make_float2(TQB_FMA(r, b.x, a.x), TQB_FMA(r, b.y, a.y));
    a = make_float2(ax, ay);
  }
}
TQB_FN void pair_hs(float2& a, float2& b) {
  const float2 t = a;
  a = make_float2(t.x + b.x, t.y + b.y);
  b = make_float2(t.x - b.x, t.y - b.y);
}

// f(a, b) on every pair of register bit J; JC >= 0: a CNOT controlled by register bit JC on this bit is folded
// in, i.e. the pair's outputs are exchanged where bit JC of e is set (compile-time: pure register renaming)
template <int R, int J, int JC, class F>
TQB_FN void for_pairs(float2 (&x)[1 << R], F f) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    float2 a = x[e], b = x[e | (1 << J)];
    f(a, b);
    const bool sw = JC >= 0 && ((e >> (JC >= 0 ? JC : 0)) & 1);
    x[e] = sw ? b : a;
    x[e | (1 << J)] = sw ? a : b;
  }
}

// scaled / CNOT-folding forward ops (KIND: TQB_T_RXS, TQB_T_RYS, TQB_T_HS, TQB_T_DIAGS with F_DJ, TQB_T_REAL)
template <int R, int J, int JC, int KIND>
TQB_FN void fast1q(float2 (&x)[1 << R], const Mat& m) {
  if constexpr (KIND == TQB_T_HS) {
    for_pairs<R, J, JC>(x, [](float2& a, float2& b) { pair_hs(a, b); });
  } else {
    const float4 q = m.q0;
    if constexpr (KIND == TQB_T_RXS) {
      const float r = q.z;
      if (q.w == 0.f) for_pairs<R, J, JC>(x, [r](float2& a, float2& b) { pair_rxs<0>(a, b, r); });
      else for_pairs<R, J, JC>(x, [r](float2& a, float2& b) { pair_rxs<1>(a, b, r); });
    } else if constexpr (KIND == TQB_T_RYS) {
      const float r = q.z;
      if (q.w == 0.f) for_pairs<R, J, JC>(x, [r](float2& a, float2& b) { pair_rys<0>(a, b, r); });
      else for_pairs<R, J, JC>(x, [r](float2& a, float2& b) { pair_rys<1>(a, b, r); });
    } else if constexpr (KIND == TQB_T_DIAGS) {
      const float2 rel = make_float2(q.x, q.y);
      for_pairs<R, J, JC>(x, [rel](float2&, float2& b) { cmul_inplace(rel, b); });
    } else {  // TQB_T_REAL
      for_pairs<R, J, JC>(x, [q](float2& a, float2& b) { pair_real(a, b, q.x, q.y, q.z, q.w); });
    }
  }
}

template <int R, int J, int KIND, bool MASKED, bool ADJ>
TQB_FN void u1q(float2 (&x)[1 << R], const Mat& m, u32 act) {
  if constexpr (KIND == TQB_T_GEN) {
    const float4 q0 = m.q0, q1 = ld4(m.p + 4);
    float2 m00 = make_float2(q0.x, q0.y), m01 = make_float2(q0.z, q0.w);
    float2 m10 = make_float2(q1.x, q1.y), m11 = make_float2(q1.z, q1.w);
    if (ADJ) {
      const float2 t = m01;
      m00 = cconj(m00); m01 = cconj(m10); m10 = cconj(t); m11 = cconj(m11);
    }
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      if (e & (1 << J)) continue;
      if (!MASKED || ((act >> e) & 1u)) pair_gen(x[e], x[e | (1 << J)], m00, m01, m10, m11);
    }
  } else if constexpr (KIND == TQB_T_REAL) {
    const float4 q = m.q0;
    const float m00 = q.x, m01 = ADJ ? q.z : q.y, m10 = ADJ ? q.y : q.z, m11 = q.w;
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      if (e & (1 << J)) continue;
      if (!MASKED || ((act >> e) & 1u)) pair_real(x[e], x[e | (1 << J)], m00, m01, m10, m11);
    }
  } else {  // TQB_T_RX
    const float4 q = m.q0;
    const float cc = q.x, s = ADJ ? -q.y : q.y;
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      if (e & (1 << J)) continue;
      if (!MASKED || ((act >> e) & 1u)) pair_rx(x[e], x[e | (1 << J)], cc, s);
    }
  }
}

template <int R, int J>
TQB_FN void perm(float2 (&x)[1 << R], u32 act) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const bool p = (act >> e) & 1u;
    const float2 a = x[e], b = x[e | (1 << J)];
    x[e] = p ? b : a;
    x[e | (1 << J)] = p ? a : b;
  }
}

// amp *= (element parity ? dB : dA) on active elements
template <int R, bool MASKED>
TQB_FN void diag(float2 (&x)[1 << R], float2 dA, float2 dB, u32 epar, u32 act) {
  if (epar == 0u) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!MASKED || ((act >> e) & 1u)) cmul_inplace(dA, x[e]);
  } else {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!MASKED || ((act >> e) & 1u)) cmul_inplace(((epar >> e) & 1u) ? dB : dA, x[e]);
  }
}

// DIAG whose element parity is bit J of e: no per-element selects
template <int R, int J, bool MASKED>
TQB_FN void diagj(float2 (&x)[1 << R], float2 dA, float2 dB, u32 act) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    if (!MASKED || ((act >> e) & 1u)) cmul_inplace((e & (1 << J)) ? dB : dA, x[e]);
}

// 4x4 on register bits (1,0): quad index q = 2*bit1 + bit0
template <int R, bool ADJ>
TQB_FN void u2q(float2 (&x)[1 << R], const Mat& mm, u32 act) {
  const float* m = mm.p;
#pragma unroll
  for (int e = 0; e < (1 << R); e += 4) {
    if ((act >> e) & 1u) {
      float2 v[4], o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = x[e + q];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* p = ADJ ? m + 2 * (4 * j + i) : m + 2 * (4 * i + j);
          const float2 mv = make_float2(p[0], ADJ ? -p[1] : p[1]);
          o[i] = cadd(o[i], cmul(mv, v[j]));
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) x[e + q] = o[q];
    }
  }
}

// --- gradient accumulators (x = psi_in, y = lam_out unless noted) -------------------------
// G[2i+j] += lam_out_i conj(psi_in_j) over active pairs of register bit J
template <int R, int J, bool MASKED>
TQB_FN void grad_full_1q(const float2 (&x)[1 << R], const float2 (&y)[1 << R], u32 act, float (&g)[8]) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if (!MASKED || ((act >> e) & 1u)) {
      const float2 a = x[e], b = x[e | (1 << J)], la = y[e], lb = y[e | (1 << J)];
      float2 t;
      t = cmulc(a, la); g[0] += t.x; g[1] += t.y;
      t = cmulc(b, la); g[2] += t.x; g[3] += t.y;
      t = cmulc(a, lb); g[4] += t.x; g[5] += t.y;
      t = cmulc(b, lb); g[6] += t.x; g[7] += t.y;
    }
  }
}

// generators X / Y of exp(-i theta P / 2) on register bit J (either side of the gate works: P commutes with it):
//   GX: w += Im(conj(psi_b) lam_a + conj(psi_a) lam_b);  GY: w += Re(conj(psi_b) lam_a - conj(psi_a) lam_b)
template <int R, int J, bool IS_Y, bool MASKED>
TQB_FN float grad_gen_1q(const float2 (&x)[1 << R], const float2 (&y)[1 << R], u32 act) {
  float w = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if (!MASKED || ((act >> e) & 1u)) {
      const float2 a = x[e], b = x[e | (1 << J)], la = y[e], lb = y[e | (1 << J)];
      if (IS_Y) {
        w = TQB_FMA(b.x, la.x, w); w = TQB_FMA(b.y, la.y, w);
        w = TQB_FMA(-a.x, lb.x, w); w = TQB_FMA(-a.y, lb.y, w);
      } else {
        w = TQB_FMA(b.x, la.y, w); w = TQB_FMA(-b.y, la.x, w);
        w = TQB_FMA(a.x, lb.y, w); w = TQB_FMA(-a.y, lb.x, w);
      }
    }
  }
  return w;
}

// w[p] += Im(conj(psi_k) lam_k) over active elements of parity p (tp: thread parity, epar: element parity)
template <int R, bool MASKED>
TQB_FN void grad_dw(const float2 (&x)[1 << R], const float2 (&y)[1 << R], u32 tp, u32 epar, u32 act, float (&w)[2]) {
  float wa = 0.f, wb = 0.f;  // wa: elements with element parity 0, wb: parity 1
  if (epar == 0u) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!MASKED || ((act >> e) & 1u)) { wa = TQB_FMA(x[e].x, y[e].y, wa); wa = TQB_FMA(-x[e].y, y[e].x, wa); }
  } else {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      if (!MASKED || ((act >> e) & 1u)) {
        const float v = TQB_FMA(x[e].x, y[e].y, -(x[e].y * y[e].x));
        if ((epar >> e) & 1u) wb += v; else wa += v;
      }
    }
  }
  w[0] = tp ? wb : wa;
  w[1] = tp ? wa : wb;
}

template <int R, int J, bool MASKED>
TQB_FN void grad_dwj(const float2 (&x)[1 << R], const float2 (&y)[1 << R], u32 tp, u32 act, float (&w)[2]) {
  float wa = 0.f, wb = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (!MASKED || ((act >> e) & 1u)) {
      if (e & (1 << J)) { wb = TQB_FMA(x[e].x, y[e].y, wb); wb = TQB_FMA(-x[e].y, y[e].x, wb); }
      else { wa = TQB_FMA(x[e].x, y[e].y, wa); wa = TQB_FMA(-x[e].y, y[e].x, wa); }
    }
  }
  w[0] = tp ? wb : wa;
  w[1] = tp ? wa : wb;
}

template <int R>
TQB_FN void grad_full_2q(const float2 (&x)[1 << R], const float2 (&y)[1 << R], u32 act, float (&g)[32]) {
#pragma unroll
  for (int e = 0; e < (1 << R); e += 4) {
    if ((act >> e) & 1u) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = cmulc(x[e + j], y[e + i]);
          g[2 * (4 * i + j)] += t.x; g[2 * (4 * i + j) + 1] += t.y;
        }
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

#define TQB_DISPATCH_J(j, CALL)                               \
  do {                                                        \
    if ((j) == 0) { constexpr int JJ = 0; CALL; }             \
    if constexpr (R > 1) { if ((j) == 1) { constexpr int JJ = 1; CALL; } } \
    if constexpr (R > 2) { if ((j) == 2) { constexpr int JJ = 2; CALL; } } \
    if constexpr (R > 3) { if ((j) == 3) { constexpr int JJ = 3; CALL; } } \
  } while (0)

// JJ from j, then CC from the folded CNOT's control register bit (CC = -1: none); j == jc never occurs
#define TQB_DISPATCH_JC(j, postx, jc, CALL)                                                          \
  TQB_DISPATCH_J(j, do {                                                                             \
    if (!(postx)) { constexpr int CC = -1; CALL; }                                                   \
    else {                                                                                           \
      if constexpr (JJ != 0) { if ((jc) == 0) { constexpr int CC = 0; CALL; } }                      \
      if constexpr (R > 1 && JJ != 1) { if ((jc) == 1) { constexpr int CC = 1; CALL; } }             \
      if constexpr (R > 2 && JJ != 2) { if ((jc) == 2) { constexpr int CC = 2; CALL; } }             \
      if constexpr (R > 3 && JJ != 3) { if ((jc) == 3) { constexpr int CC = 3; CALL; } }             \
    }                                                                                                \
  } while (0))

// forward op, in place on the thread's register array.  Dispatch is an if-chain ordered by frequency: a switch /
// jump table measured 10 % slower (profiles/r01_kernel_variants.txt)
template <int R, class Exec>
TQB_FN void apply_op_fwd_inplace(const BlockCtx& c, float2 (&x)[1 << R], const int4 h, const Mat& m, const int32_t* op, u32 base) {
  const u32 w0 = (u32)h.x;
  if (w0 & TQB_F_OCTRL) {
    if ((c.tile_ext & (u32)op[TQB_OP_CM_O]) != (u32)op[TQB_OP_CV_O]) return;
  }
  const u32 type = TQB_OP_TYPE(w0);
  const int j = TQB_OP_J(w0);
  const bool masked = w0 & (TQB_F_TCTRL | TQB_F_RCTRL);
  u32 act = 0xffffu;
  if (masked) {
    act = (u32)h.z & 0xffffu;
    if ((w0 & TQB_F_TCTRL) && (base & (u32)op[TQB_OP_CM_T]) != (u32)op[TQB_OP_CV_T]) act = 0u;
  }
  if (type >= TQB_T_RXS || (w0 & TQB_F_POSTX)) {
    // scaled forms and ops with a folded CNOT (programs without a backward)
    const bool postx = w0 & TQB_F_POSTX;
    const int jc = TQB_OP_JC(w0);
    if (TQB_HAS(TQB_T_RXS) && type == TQB_T_RXS) {
      TQB_DISPATCH_JC(j, postx, jc, (fast1q<R, JJ, CC, TQB_T_RXS>(x, m)));
    } else if (TQB_HAS(TQB_T_RYS) && type == TQB_T_RYS) {
      TQB_DISPATCH_JC(j, postx, jc, (fast1q<R, JJ, CC, TQB_T_RYS>(x, m)));
    } else if (TQB_HAS(TQB_T_DIAGS) && type == TQB_T_DIAGS) {
      if (w0 & TQB_F_DJ) {
        TQB_DISPATCH_JC(j, postx, jc, (fast1q<R, JJ, CC, TQB_T_DIAGS>(x, m)));
      } else {
        // the bit lives outside the registers: the thread's 2^R amplitudes all share its value
        const u32 tp = (u32)(__popc(base & (u32)h.w) + __popc(c.tile_ext & (u32)op[TQB_OP_PM_O])) & 1u;
        if (tp) {
          const float4 q = m.q0;
          const float2 rel = make_float2(q.x, q.y);
#pragma unroll
          for (int e = 0; e < (1 << R); ++e) cmul_inplace(rel, x[e]);
        }
      }
    } else if (TQB_HAS(TQB_T_HS) && type == TQB_T_HS) {
      TQB_DISPATCH_JC(j, postx, jc, (fast1q<R, JJ, CC, TQB_T_HS>(x, m)));
    } else if (TQB_HAS(TQB_T_RXS)) {  // TQB_T_REAL with a folded CNOT (forward-only programs)
      TQB_DISPATCH_JC(j, postx, jc, (fast1q<R, JJ, CC, TQB_T_REAL>(x, m)));
    }
    return;
  }
  if (TQB_HAS(TQB_T_REAL) && type == TQB_T_REAL) {
    TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_REAL, false, false>(x, m, act)));
  } else if (TQB_HAS(TQB_T_RX) && type == TQB_T_RX) {
    TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_RX, false, false>(x, m, act)));
  } else if (TQB_HAS(TQB_T_DIAG) && type == TQB_T_DIAG) {
    const float4 q = m.q0;
    const u32 tp = (u32)(__popc(base & (u32)h.w) + __popc(c.tile_ext & (u32)op[TQB_OP_PM_O])) & 1u;
    const float2 d0 = make_float2(q.x, q.y), d1 = make_float2(q.z, q.w);
    const float2 dA = tp ? d1 : d0, dB = tp ? d0 : d1;
    const u32 epar = (u32)h.z >> 16;
    if (w0 & TQB_F_DJ) {
      if (masked) { if (act) TQB_DISPATCH_J(j, (diagj<R, JJ, true>(x, dA, dB, act))); }
      else TQB_DISPATCH_J(j, (diagj<R, JJ, false>(x, dA, dB, act)));
    } else if (masked) { if (act) diag<R, true>(x, dA, dB, epar, act); }
    else diag<R, false>(x, dA, dB, epar, act);
  } else if (TQB_HAS(TQB_T_PERM) && type == TQB_T_PERM) {
    if (act) TQB_DISPATCH_J(j, (perm<R, JJ>(x, act)));
  } else if (TQB_HAS(TQB_T_GEN) && type == TQB_T_GEN) {
    if (masked) { if (act) TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_GEN, true, false>(x, m, act))); }
    else TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_GEN, false, false>(x, m, act)));
  } else if (TQB_HAS(TQB_T_U2Q)) {  // TQB_T_U2Q
    if constexpr (R >= 2) { if (act) u2q<R, false>(x, m, act); }
  }
}

// backward of one op: x holds psi_out (psi_in when NOUNAPPLY), y holds lam_out; leaves psi_in / lam_in.
// The masked (controlled) and unmasked variants are split per op TYPE, never around the whole dispatcher: two
// instantiations of the full dispatcher made the compiler permute the 64 live amplitude registers at every op
// (ncu: 27 % of the adjoint kernel's instructions were MOVs, 19 % on the line of that split).
template <int R, bool MASKED>
TQB_FN void bwd_gen(float2 (&x)[1 << R], float2 (&y)[1 << R], const Mat& m, int j, u32 act, u32 gkind, bool unapply,
                    float* spart, int pstride, int gslot) {
  float gacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float w = 0.f;
  if (!MASKED || act) {
    if (unapply) TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_GEN, MASKED, true>(x, m, act)));
    if (gkind == TQB_G_FULL) TQB_DISPATCH_J(j, (grad_full_1q<R, JJ, MASKED>(x, y, act, gacc)));
    TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_GEN, MASKED, true>(y, m, act)));
  }
  if (gkind == TQB_G_X) TQB_DISPATCH_J(j, (w = grad_gen_1q<R, JJ, false, MASKED>(x, y, act)));
  if (gkind == TQB_G_Y) TQB_DISPATCH_J(j, (w = grad_gen_1q<R, JJ, true, MASKED>(x, y, act)));
  if (gkind == TQB_G_FULL) stash<8>(spart, pstride, gslot, gacc);
  else if (gkind != TQB_G_NONE) { float w1[1] = {w}; stash<1>(spart, pstride, gslot, w1); }
}

template <int R, bool MASKED>
TQB_FN void bwd_diag(float2 (&x)[1 << R], float2 (&y)[1 << R], float2 dA, float2 dB, u32 w0, int j, u32 epar, u32 tp,
                     u32 act, u32 gkind, bool unapply, float* spart, int pstride, int gslot) {
  if (w0 & TQB_F_DJ) {
    if (!MASKED || act) {
      if (unapply) TQB_DISPATCH_J(j, (diagj<R, JJ, MASKED>(x, dA, dB, act)));
      TQB_DISPATCH_J(j, (diagj<R, JJ, MASKED>(y, dA, dB, act)));
    }
    if (gkind == TQB_G_DW) {
      float w2[2] = {0.f, 0.f};
      TQB_DISPATCH_J(j, (grad_dwj<R, JJ, MASKED>(x, y, tp, act, w2)));
      stash<2>(spart, pstride, gslot, w2);
    }
  } else {
    if (!MASKED || act) {
      if (unapply) diag<R, MASKED>(x, dA, dB, epar, act);
      diag<R, MASKED>(y, dA, dB, epar, act);
    }
    if (gkind == TQB_G_DW) {
      float w2[2];
      grad_dw<R, MASKED>(x, y, tp, epar, act, w2);
      stash<2>(spart, pstride, gslot, w2);
    }
  }
}

template <int R, class Exec>
TQB_FN void apply_op_bwd(Exec& ex, const BlockCtx& c, float2 (&x)[1 << R], float2 (&y)[1 << R], const int4 h, const Mat& m,
                         const int32_t* op, u32 base, float* spart, int pstride) {
  const u32 w0 = (u32)h.x;
  if (w0 & TQB_F_OCTRL) {
    if ((c.tile_ext & (u32)op[TQB_OP_CM_O]) != (u32)op[TQB_OP_CV_O]) return;
  }
  const bool masked = w0 & (TQB_F_TCTRL | TQB_F_RCTRL);
  u32 act = 0xffffu;
  if (masked) {
    act = (u32)h.z & 0xffffu;
    if ((w0 & TQB_F_TCTRL) && (base & (u32)op[TQB_OP_CM_T]) != (u32)op[TQB_OP_CV_T]) act = 0u;
  }
  const u32 type = TQB_OP_TYPE(w0);
  const int j = TQB_OP_J(w0);
  const u32 gkind = TQB_OP_GKIND(w0);
  const bool unapply = !(w0 & TQB_F_NOUNAPPLY);
  const int gslot = op[TQB_OP_GSLOT];
  if (type == TQB_T_REAL) {          // never controlled (planner): ry, h, ...
    TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_REAL, false, true>(x, m, act), u1q<R, JJ, TQB_T_REAL, false, true>(y, m, act)));
    if (gkind == TQB_G_Y) {
      float w1[1] = {0.f};
      TQB_DISPATCH_J(j, (w1[0] = grad_gen_1q<R, JJ, true, false>(x, y, act)));
      stash<1>(spart, pstride, gslot, w1);
    }
  } else if (type == TQB_T_RX) {     // never controlled
    TQB_DISPATCH_J(j, (u1q<R, JJ, TQB_T_RX, false, true>(x, m, act), u1q<R, JJ, TQB_T_RX, false, true>(y, m, act)));
    if (gkind == TQB_G_X) {
      float w1[1] = {0.f};
      TQB_DISPATCH_J(j, (w1[0] = grad_gen_1q<R, JJ, false, false>(x, y, act)));
      stash<1>(spart, pstride, gslot, w1);
    }
  } else if (type == TQB_T_DIAG) {
    const float4 q = m.q0;
    const u32 tp = (u32)(__popc(base & (u32)h.w) + __popc(c.tile_ext & (u32)op[TQB_OP_PM_O])) & 1u;
    const float2 d0 = make_float2(q.x, -q.y), d1 = make_float2(q.z, -q.w);  // conjugated
    const float2 dA = tp ? d1 : d0, dB = tp ? d0 : d1;
    const u32 epar = (u32)h.z >> 16;
    if (masked) bwd_diag<R, true>(x, y, dA, dB, w0, j, epar, tp, act, gkind, unapply, spart, pstride, gslot);
    else bwd_diag<R, false>(x, y, dA, dB, w0, j, epar, tp, act, gkind, unapply, spart, pstride, gslot);
  } else if (type == TQB_T_PERM) {
    if (act) TQB_DISPATCH_J(j, (perm<R, JJ>(x, act), perm<R, JJ>(y, act)));
  } else if (TQB_HAS(TQB_T_GEN) && type == TQB_T_GEN) {
    if (masked) bwd_gen<R, true>(x, y, m, j, act, gkind, unapply, spart, pstride, gslot);
    else bwd_gen<R, false>(x, y, m, j, act, gkind, unapply, spart, pstride, gslot);
  } else if (TQB_HAS(TQB_T_U2Q)) {  // TQB_T_U2Q
    if constexpr (R >= 2) {
      float g32[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) g32[i] = 0.f;
      if (act) {
        if (unapply) u2q<R, true>(x, m, act);
        if (gkind == TQB_G_FULL) grad_full_2q<R>(x, y, act, g32);
        u2q<R, true>(y, m, act);
      }
      if (gkind == TQB_G_FULL) stash<32>(spart, pstride, gslot, g32);
    }
  }
}

// TQB_T_LADDER: the phase contributed by the ladder bits in the outer (non-tile) index is the same for the whole
// tile: one thread per ladder op multiplies it up and parks it in the op's staged record (TQB_OP_LPH)
TQB_FN void ladder_outer(const BlockCtx& c, int tid) {
  const int ng = c.s_prog[TQB_S_NGROUPS], nops = c.s_prog[TQB_S_NOPS];
  int32_t* ops = c.s_prog + TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS;
  for (int k = tid; k < nops; k += c.nthreads) {
    int32_t* op = ops + k * TQB_OP_WORDS;
    if (TQB_OP_TYPE((u32)op[TQB_OP_W0]) != TQB_T_LADDER) continue;
    const float* tab = c.s_msh + op[TQB_OP_SLOT];
    float2 ph = make_float2(1.f, 0.f);
    u32 mo = c.tile_ext & (u32)op[TQB_OP_PM_O];
    while (mo) {
      const int b = __ffs((int)mo) - 1;
      mo &= mo - 1u;
      ph = cmul(ph, *reinterpret_cast<const float2*>(tab + 2 * (c.T + b)));
    }
    *reinterpret_cast<float2*>(op + TQB_OP_LPH) = ph;
  }
}

// A ladder group (group record word 7 == 1) holds only TQB_T_LADDER ops and no register bits: it is applied
// straight on the staged tile, 16 amplitudes (tile-local bits 0-3) per thread and item.  It lives outside the
// register interpreter on purpose: with its data-dependent loops inside apply_op_fwd_inplace every other op of the
// forward kernel ran 35-40 % slower (profiles/r01_v3_notes.txt).
// Separate function (not inlined): with this body inlined into the pass kernel, ptxas produced ~20 % slower code
// for the whole kernel (every op type), although the group is never executed by those programs.
TQB_COLD void ladder_group(const int32_t* s_prog, float2* s_psi, const float* s_msh, int T, int tps, u32 tile_ext, bool pair,
                           int tid, int g) {
  struct { const int32_t* s_prog; float2* s_psi; const float* s_msh; int T, tps; u32 tile_ext; } c = {s_prog, s_psi, s_msh, T, tps, tile_ext};
  const int ng = c.s_prog[TQB_S_NGROUPS];
  const int32_t* gh = c.s_prog + TQB_SEG_HDR_WORDS + g * TQB_GROUP_WORDS;
  const int nops = gh[4];
  const int32_t* ops = c.s_prog + TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS + gh[5] * TQB_OP_WORDS;
  const int slot = tid / c.tps, t = tid % c.tps;
  float2* sps = c.s_psi + ((size_t)slot << c.T);
  const u32 tile = 1u << c.T;
  const u32 ne = tile < 16u ? tile : 16u;         // amplitudes per item
  const u32 items = tile < 16u ? 1u : tile >> 4;
  for (int k = 0; k < nops; ++k) {
    const int32_t* op = ops + k * TQB_OP_WORDS;
    const u32 w0 = (u32)op[TQB_OP_W0];
    if ((w0 & TQB_F_OCTRL) && (c.tile_ext & (u32)op[TQB_OP_CM_O]) != (u32)op[TQB_OP_CV_O]) continue;
    const float* tab = c.s_msh + op[TQB_OP_SLOT];
    const float2 lph = *reinterpret_cast<const float2*>(op + TQB_OP_LPH);
    const u32 pm = (u32)op[TQB_OP_PM_T];
    const u32 cm = (w0 & TQB_F_TCTRL) ? (u32)op[TQB_OP_CM_T] : 0u, cv = (w0 & TQB_F_TCTRL) ? (u32)op[TQB_OP_CV_T] : 0u;
    // phases of the 16 patterns of tile-local bits 0-3
    float2 P[16];
    P[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float2 tj = make_float2(1.f, 0.f);
      if ((pm >> jj) & 1u) tj = *reinterpret_cast<const float2*>(tab + 2 * jj);
#pragma unroll
      for (int e = (1 << jj); e < (2 << jj); ++e) P[e] = cmul(P[e ^ (1 << jj)], tj);
    }
    // phase of the ladder bits >= 4 that are set in `hi`, times the tile's outer phase
    auto hi_phase = [&](u32 hi) {
      float2 ph = lph;
      u32 mt = hi & pm;
      while (mt) {
        const int b = __ffs((int)mt) - 1;
        mt &= mt - 1u;
        ph = cmul(ph, *reinterpret_cast<const float2*>(tab + 2 * b));
      }
      return ph;
    };
    if (w0 & TQB_F_PREH) {
      // H on the control bit c, then the ladder: (a, b) -> (a + b, (a - b) * phase(b's index)) on every pair of bit c
      const u32 pc = TQB_OP_PC(w0), cb = 1u << pc;
      if (pc < 4u) {
        // (compile-time bit position: a run-time index into P[] would push the array into local memory)
        auto low_case = [&](auto pcc) {
          constexpr u32 CB = 1u << decltype(pcc)::value;
          for (u32 item = (u32)t; item < items; item += (u32)c.tps) {
            const u32 hi = item << 4;
            const float2 ph = hi_phase(hi);
#pragma unroll
            for (u32 e = 0; e < 16u; ++e) {
              if (!(e & CB) && (e | CB) < ne) {
                const u32 a0 = swz(hi | e, pair), a1 = swz(hi | e | CB, pair);
                const float2 A = sps[a0], B = sps[a1];
                sps[a0] = make_float2(A.x + B.x, A.y + B.y);
                sps[a1] = cmul(cmul(ph, P[e | CB]), make_float2(A.x - B.x, A.y - B.y));
              }
            }
          }
        };
        if (pc == 0u) low_case(std::integral_constant<int, 0>{});
        else if (pc == 1u) low_case(std::integral_constant<int, 1>{});
        else if (pc == 2u) low_case(std::integral_constant<int, 2>{});
        else low_case(std::integral_constant<int, 3>{});
      } else {
        const u32 lowm = (1u << (pc - 4u)) - 1u;
        for (u32 it = (u32)t; it < (items >> 1); it += (u32)c.tps) {
          const u32 hi0 = (((it & ~lowm) << 1) | (it & lowm)) << 4, hi1 = hi0 | cb;
          const float2 ph = hi_phase(hi1);
#pragma unroll
          for (u32 e = 0; e < 16u; ++e) {
            if (e < ne) {
              const u32 a0 = swz(hi0 | e, pair), a1 = swz(hi1 | e, pair);
              const float2 A = sps[a0], B = sps[a1];
              sps[a0] = make_float2(A.x + B.x, A.y + B.y);
              sps[a1] = cmul(cmul(ph, P[e]), make_float2(A.x - B.x, A.y - B.y));
            }
          }
        }
      }
      continue;
    }
    for (u32 item = (u32)t; item < items; item += (u32)c.tps) {
      const u32 hi = item << 4;
      if ((hi & cm & ~15u) != (cv & ~15u)) continue;
      const float2 ph = hi_phase(hi);
#pragma unroll
      for (u32 e = 0; e < 16u; ++e) {
        if (e < ne && (e & cm) == (cv & 15u)) {
          const u32 a = swz(hi | e, pair);
          sps[a] = cmul(cmul(ph, P[e]), sps[a]);
        }
      }
    }
  }
}

template <int R, bool BWD, class Exec>
TQB_FN void run_group(Exec& ex, const BlockCtx& c, int tid, int g) {
  const int ng = c.s_prog[TQB_S_NGROUPS];
  const int32_t* gh = c.s_prog + TQB_SEG_HDR_WORDS + g * TQB_GROUP_WORDS;
  const int slot = tid / c.tps, t = tid % c.tps;
  u32 sp[R], zp[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { sp[j] = (u32)gh[j]; zp[j] = swz(1u << sp[j], c.multi_tile); }
  sort_positions<R>(sp);
  const int nops = gh[4];
  const int32_t* ops = c.s_prog + TQB_SEG_HDR_WORDS + ng * TQB_GROUP_WORDS + gh[5] * TQB_OP_WORDS;
  const int items = 1 << (c.T - R);
  float2* sps = c.s_psi + ((size_t)slot << c.T);
  float2* sla = BWD ? c.s_lam + ((size_t)slot << c.T) : nullptr;
  const float* smps = c.s_mps + slot * c.ps_cap;
  float* spart = BWD ? c.s_part + tid : nullptr;
  const int pstride = c.nthreads + 1;

  for (int item = t; item < items; item += c.tps) {
    const u32 base = insert_zeros<R>((u32)item, sp);
    const u32 sb = swz(base, c.multi_tile);
    float2 x[1 << R];
    float2 y[1 << R];
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      u32 a = sb;
#pragma unroll
      for (int j = 0; j < R; ++j) if (e & (1 << j)) a ^= zp[j];
      x[e] = sps[a];
      if constexpr (BWD) y[e] = sla[a];
    }
    // op header (w0, slot, emask, pm_t) in one 128-bit shared read, then the first four matrix floats.
    // (Fetching them one / two ops ahead was tried: the register rotation costs more ALU-pipe MOVs than the
    // latency it hides - profiles/r01_v3_notes.txt.)
    for (int i = 0; i < nops; ++i) {
      const int k = BWD ? nops - 1 - i : i;
      const int32_t* op = ops + k * TQB_OP_WORDS;
      const int4 h = *reinterpret_cast<const int4*>(op);
      Mat m;
      m.p = ((h.x & TQB_F_PS) ? smps : c.s_msh) + h.y;
      m.q0 = ld4(m.p);
      if constexpr (BWD) apply_op_bwd<R, Exec>(ex, c, x, y, h, m, op, base, spart, pstride);
      else apply_op_fwd_inplace<R, Exec>(c, x, h, m, op, base);
    }
#pragma unroll
    for (int e = 0; e < (1 << R); ++e) {
      u32 a = sb;
#pragma unroll
      for (int j = 0; j < R; ++j) if (e & (1 << j)) a ^= zp[j];
      sps[a] = x[e];
      if constexpr (BWD) sla[a] = y[e];
    }
  }
}

// all segments / groups of the pass on the tile staged in c.s_psi (and c.s_lam)
template <int R, bool BWD, class Exec>
TQB_FN void run_tile(Exec& ex, const BlockCtx& c, bool shared_staged) {
  const int nseg = c.prog[TQB_H_NSEG];
  for (int si = 0; si < nseg; ++si) {
    const int s = BWD ? nseg - 1 - si : si;
    if (!shared_staged) {
      const int soff = c.prog[TQB_H_SEGOFF + s];
      ex.par([&](int tid) { stage_segment_shared(c, tid, soff); });
    }
    if (c.s_prog[TQB_S_PS_LEN] > 0 || (BWD && c.s_prog[TQB_S_G_LEN] > 0))
      ex.par([&](int tid) { stage_segment_sample<BWD>(c, tid); });
    const int ng = c.s_prog[TQB_S_NGROUPS];
    if (!BWD && c.s_prog[TQB_SEG_HDR_WORDS + 6] > 0) ex.par([&](int tid) { ladder_outer(c, tid); });
    for (int gi = 0; gi < ng; ++gi) {
      const int g = BWD ? ng - 1 - gi : gi;
      if (!BWD && c.s_prog[TQB_SEG_HDR_WORDS + g * TQB_GROUP_WORDS + 7] == 1) ex.par([&](int tid) { ladder_group(c.s_prog, c.s_psi, c.s_msh, c.T, c.tps, c.tile_ext, c.multi_tile, tid, g); });
      else ex.par([&](int tid) { run_group<R, BWD>(ex, c, tid, g); });
    }
    if (BWD && c.s_prog[TQB_S_G_LEN] > 0) {
      ex.par([&](int tid) { reduce_partials(c, tid); });
      ex.par([&](int tid) { flush_grads(ex, c, tid); });
    }
  }
}

// persistent loop of one CTA over its work items (a work item = S samples of a single-tile state,
// or one tile of one sample), with the next item's loads in flight while the current one computes
// (c is taken by mutable reference on purpose: nvcc 12.9 mis-lowers a BlockCtx copy made here after
// inlining into the kernel - the copy ends up aliased with the kernel's parameter space)
template <int R, bool BWD, class Exec>
TQB_FN void run_cta(Exec& ex, BlockCtx& c, int n, u32 ntiles_log2, u32 ext_hi, long long item0, long long item_step,
                    long long n_items, int nbuf, size_t buf_stride /* float2 elements between buffers */) {
  float2* const psi_buf = c.s_psi;
  float2* const lam_buf = c.s_lam;
  const bool shared_once = c.prog[TQB_H_NSEG] == 1;
  const bool streamed = ntiles_log2 > 0;
  ex.par([&](int tid) {
    build_tables(c, tid, n, streamed);
    if (shared_once) stage_segment_shared(c, tid, c.prog[TQB_H_SEGOFF]);
  });
  auto locate = [&](long long item) {
    if (streamed) {
      const u32 tile_id = (u32)item & ((1u << ntiles_log2) - 1u);
      c.sample0 = (int)(item >> ntiles_log2);
      c.tile_ext = tile_id | ext_hi;
      c.outer_off = c.s_olut[tile_id & 255u] + c.s_olut[256u + ((tile_id >> 8) & 255u)] + c.s_olut[512u + (tile_id >> 16)];
    } else {
      c.sample0 = (int)(item * c.S);
      c.tile_ext = ext_hi;
      c.outer_off = 0ull;
    }
  };
  auto issue = [&](long long item, int buf) {
    locate(item);
    ex.par_nosync([&](int tid) {
      load_tile_async(ex, c, tid, c.psi_src, psi_buf + buf * buf_stride);
      if (BWD) load_tile_async(ex, c, tid, c.lam_src, lam_buf + buf * buf_stride);
      ex.async_commit();
    });
  };
  if (item0 < n_items) issue(item0, 0);
  int buf = 0;
  for (long long item = item0; item < n_items; item += item_step) {
    const long long next = item + item_step;
    if (nbuf > 1 && next < n_items) {
      issue(next, buf ^ 1);
      ex.par([&](int) { ex.async_wait_prev(); });
    } else {
      ex.par([&](int) { ex.async_wait_all(); });
    }
    locate(item);
    c.s_psi = psi_buf + buf * buf_stride;
    c.s_lam = lam_buf + buf * buf_stride;
    run_tile<R, BWD>(ex, c, shared_once);
    ex.par([&](int tid) {
      store_tile(c, tid, c.psi, c.s_psi);
      if (BWD) store_tile(c, tid, c.lam, c.s_lam);
    });
    if (nbuf > 1) buf ^= 1;
    else if (next < n_items) issue(next, 0);
  }
}

// shared-memory carve-up (bytes)
struct SmemLayout { size_t psi, lam, msh, mps, grad, part, prog, scat, olut, total, buf_stride; };
inline SmemLayout smem_layout(int T, int L, int S, int tps, bool bwd, int nbuf, int sh_max, int ps_max, int g_max, bool streamed) {
  SmemLayout l;
  size_t off = 0;
  const size_t tile = ((size_t)S << T) * sizeof(float2);
  l.buf_stride = tile / sizeof(float2);
  l.psi = off; off += tile * nbuf;
  l.lam = off; if (bwd) off += tile * nbuf;
  l.msh = off; off += (size_t)((sh_max + 3) & ~3) * sizeof(float);
  l.mps = off; off += (size_t)S * ps_max * sizeof(float);
  l.grad = off; if (bwd) off += (size_t)S * g_max * sizeof(float);
  l.part = off; if (bwd) off += (size_t)g_max * (S * tps + 1) * sizeof(float);
  off = (off + 15) & ~(size_t)15;  // op headers are read as 128-bit words
  l.prog = off; off += (size_t)TQB_SEG_MAX_WORDS * sizeof(int32_t);
  off = (off + 15) & ~(size_t)15;
  l.scat = off; off += ((size_t)1 << (T - L)) * sizeof(u64);
  l.olut = off; if (streamed) off += 3 * 256 * sizeof(u64);
  l.total = (off + 15) & ~(size_t)15;
  return l;
}

// launch geometry shared by the CUDA launcher and the emulation
struct PassGeom { int S, tps, nbuf; u32 ntiles_log2; long long n_items; SmemLayout lay; };
inline bool pass_geometry(int n, long long bsz, int T, int R, int L, bool bwd, int sh_max, int ps_max, int g_max,
                          size_t smem_budget, PassGeom& g) {
  const int items = 1 << (T - R);
  g.tps = items < 256 ? items : 256;
  ps_max = (ps_max + 3) & ~3;
  g_max = (g_max + 3) & ~3;
  if (n == T) {
    g.ntiles_log2 = 0;
    g.nbuf = 1;
    int S = g.tps >= 128 ? 1 : 128 / g.tps;
    while (S > 1 && smem_layout(T, L, S, g.tps, bwd, 1, sh_max, ps_max, g_max, false).total > smem_budget) S >>= 1;
    if (S * g.tps < 32) S = 32 / g.tps;  // whole warps for the shuffle reductions
    g.S = S;
    g.n_items = (bsz + S - 1) / S;
  } else {
    g.S = 1;
    g.ntiles_log2 = (u32)(n - T);
    g.n_items = bsz << (n - T);
    g.nbuf = 2;
    if (smem_layout(T, L, 1, g.tps, bwd, 2, sh_max, ps_max, g_max, true).total > smem_budget) g.nbuf = 1;
  }
  g.lay = smem_layout(T, L, g.S, g.tps, bwd, g.nbuf, sh_max, ps_max, g_max, n != T);
  return true;
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
  // evaluated in double and rounded once: the device's float sincosf is within 1-2 ulp but its s*s + c*c is biased,
  // which showed as a 2e-5 norm drift after 4 000 gates (tools/depth_parity.py); the builder handles a few thousand
  // angles per step, so the double evaluation is free
  double sd, cd;
#if defined(__CUDA_ARCH__)
  sincos((double)a, &sd, &cd);
#else
  sd = sin((double)a); cd = cos((double)a);
#endif
  s = (float)sd; c = (float)cd;
}
// double: the block-unitary builder (tqb_block_build.cuh) instantiates build_gate<double>
TQB_FN double s_const(double, float c) { return (double)c; }
TQB_FN void s_sincos(double a, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(a, &s, &c);
#else
  s = sin(a); c = cos(a);
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

// scaled rotation (c, s) -> packed (c, s, r, mode) and the factor left out of the per-amplitude arithmetic
TQB_FN void pack_scaled_rot(int type, float cc, float ss, float* out, float2& f) {
  const bool m1 = fabsf(cc) < fabsf(ss);
  out[0] = cc; out[1] = ss; out[2] = m1 ? cc / ss : ss / cc; out[3] = m1 ? 1.f : 0.f;
  if (type == TQB_T_RXS) f = m1 ? make_float2(0.f, -ss) : make_float2(cc, 0.f);
  else f = m1 ? make_float2(ss, 0.f) : make_float2(cc, 0.f);
}
// one-bit diagonal (d0, d1) -> packed (rel, d0) with rel = d1 / d0 (|d0| = 1 for every catalogue gate)
TQB_FN void pack_scaled_diag(float2 d0, float2 d1, float* out, float2& f) {
  const float n2 = d0.x * d0.x + d0.y * d0.y;
  const float2 rel = make_float2((d1.x * d0.x + d1.y * d0.y) / n2, (d1.y * d0.x - d1.x * d0.y) / n2);
  out[0] = rel.x; out[1] = rel.y; out[2] = d0.x; out[3] = d0.y;
  f = d0;
}

// one (op, sample): params -> applied matrix (inverse folded in), packed for the op type, written to
// the shared row (b ignored) or the sample's row
// matrix floats of one parametric record and, for the scaled forms, the factor they leave out (returns true then)
TQB_FN bool build_param(const int32_t* rec, long long b, const float* params_sh, const float* params_ps, int ps_stride,
                        float* out, float2& f);

TQB_FN void build_one(const int32_t* rec, long long b, const float* params_sh, const float* params_ps, int ps_stride,
                      const float* consts, float* row, const int32_t* table) {
  const int kind = rec[TQB_B_KIND];
  float* out = row + rec[TQB_B_MOFF];
  if (kind == TQB_K_EXTERN) return;
  if (kind == TQB_K_CONST) {
    for (int i = 0; i < rec[TQB_B_NFLOAT]; ++i) out[i] = consts[rec[TQB_B_COFF] + i];
    return;
  }
  if (kind == TQB_K_SCALE) {
    // constant part (folded by the planner) times the factors of the scaled parametric records of this pass
    float2 acc = make_float2(consts[rec[TQB_B_POFF]], consts[rec[TQB_B_POFF] + 1]);
    for (int r = rec[TQB_B_COFF]; r < rec[TQB_B_COFF] + rec[TQB_B_NFLOAT]; ++r) {
      const int32_t* q = table + r * TQB_B_WORDS;
      const int t = q[TQB_B_TYPE];
      if (q[TQB_B_KIND] < TQB_K_RX || q[TQB_B_KIND] >= TQB_K_SCALE || (t != TQB_T_RXS && t != TQB_T_RYS && t != TQB_T_DIAGS))
        continue;
      float tmp[4];
      float2 f;
      if (build_param(q, b, params_sh, params_ps, ps_stride, tmp, f)) acc = cmul(acc, f);
    }
    out[0] = acc.x; out[1] = acc.y; out[2] = 0.f; out[3] = 0.f;
    return;
  }
  float2 f;
  build_param(rec, b, params_sh, params_ps, ps_stride, out, f);
}

TQB_FN bool build_param(const int32_t* rec, long long b, const float* params_sh, const float* params_ps, int ps_stride,
                        float* out, float2& f) {
  const int kind = rec[TQB_B_KIND];
  const bool inv = rec[TQB_B_FLAGS] & TQB_BF_INVERSE;
  const float* p = (rec[TQB_B_FLAGS] & TQB_BF_PER_SAMPLE) ? params_ps + b * ps_stride + rec[TQB_B_POFF]
                                                          : params_sh + rec[TQB_B_POFF];
  float th[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < rec[TQB_B_NPARAM]; ++i) th[i] = p[i];
  Cx<float> u[16];
  const int nent = build_gate<float>(kind, th, u);
  const int d = kind_dim(nent);
  const int type = rec[TQB_B_TYPE];
  if (d == 0) {
    if (type == TQB_T_LADDER) {  // one rung: the phase applied where both bits are set
      out[0] = u[1].re; out[1] = inv ? -u[1].im : u[1].im;
      return false;
    }
    if (type == TQB_T_DIAGS) {
      pack_scaled_diag(make_float2(u[0].re, inv ? -u[0].im : u[0].im), make_float2(u[1].re, inv ? -u[1].im : u[1].im), out, f);
      return true;
    }
    for (int i = 0; i < 2; ++i) { out[2 * i] = u[i].re; out[2 * i + 1] = inv ? -u[i].im : u[i].im; }
    return false;
  }
  Cx<float> a[16];  // applied matrix
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      const Cx<float> v = inv ? u[j * d + i] : u[i * d + j];
      a[i * d + j] = Cx<float>{v.re, inv ? -v.im : v.im};
    }
  if (type == TQB_T_REAL) {
    out[0] = a[0].re; out[1] = a[1].re; out[2] = a[2].re; out[3] = a[3].re;
  } else if (type == TQB_T_RX) {
    out[0] = a[0].re; out[1] = -a[1].im; out[2] = 0.f; out[3] = 0.f;
  } else if (type == TQB_T_RXS) {
    pack_scaled_rot(type, a[0].re, -a[1].im, out, f);
    return true;
  } else if (type == TQB_T_RYS) {
    pack_scaled_rot(type, a[0].re, a[2].re, out, f);
    return true;
  } else {
    for (int i = 0; i < d * d; ++i) { out[2 * i] = a[i].re; out[2 * i + 1] = a[i].im; }
  }
  return false;
}

// one (op, sample): dL/dtheta from the op's gradient accumulators
//   FULL: Re sum conj(dA_ij/dtheta_p) G_ij with A the applied matrix
//   X / Y: A = exp(-+ i theta P / 2)  ->  -+ w / 2
//   DW:    A_k = exp(i phi_p(theta)) -> sum_p phi_p' w_p
TQB_FN void build_one_bwd(const int32_t* rec, long long b, const float* params_sh, const float* params_ps,
                          int ps_stride, const float* grow, float* gsh_row, float* gps_row) {
  const int kind = rec[TQB_B_KIND];
  const int gkind = rec[TQB_B_GKIND];
  if (kind == TQB_K_EXTERN || kind == TQB_K_CONST || gkind == TQB_G_NONE) return;
  const bool inv = rec[TQB_B_FLAGS] & TQB_BF_INVERSE;
  const bool ps = rec[TQB_B_FLAGS] & TQB_BF_PER_SAMPLE;
  const float* p = ps ? params_ps + b * ps_stride + rec[TQB_B_POFF] : params_sh + rec[TQB_B_POFF];
  const float* G = grow + rec[TQB_B_GOFF];
  float* dst = (ps ? gps_row : gsh_row) + rec[TQB_B_POFF];
  const int np = rec[TQB_B_NPARAM];
  if (gkind == TQB_G_X || gkind == TQB_G_Y) {
    dst[0] = (inv ? 0.5f : -0.5f) * G[0];
    return;
  }
  if (gkind == TQB_G_DW) {
    float v;
    if (kind == TQB_K_RZ) v = 0.5f * (G[1] - G[0]);
    else if (kind == TQB_K_PHASE) v = G[1];   // d0 = 1 carries no phase
    else v = G[0] + G[1];                     // global phase
    dst[0] = inv ? -v : v;
    return;
  }
  for (int q = 0; q < np; ++q) {
    Dual th[4];
    for (int i = 0; i < 4; ++i) th[i] = Dual{i < np ? p[i] : 0.f, i == q ? 1.f : 0.f};
    Cx<Dual> u[16];
    const int nent = build_gate<Dual>(kind, th, u);
    const int d = kind_dim(nent);
    float acc = 0.f;
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        // applied A_ij = U_ij, or conj(U_ji) when inverted
        const Cx<Dual> v = inv ? u[j * d + i] : u[i * d + j];
        acc += v.re.d * G[2 * (i * d + j)] + (inv ? -v.im.d : v.im.d) * G[2 * (i * d + j) + 1];
      }
    dst[q] = acc;
  }
}

}  // namespace tqb
