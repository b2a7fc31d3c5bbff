// TEST INFRASTRUCTURE ONLY — host build of the gate-pass kernel body (csrc/tqb_core.cuh) behind
// the same C ABI as libtqb200.so, operating on HOST memory with the CUDA block/thread structure
// emulated sequentially.  It lets the CPU test-suite (-m "not gpu") drive the host planner, the
// autograd wiring and the program wire format end to end without a GPU.  The product package
// never loads this library (torchquantum_b200/_native.py only loads the CUDA build and raises if
// it is missing); tests inject it explicitly.
//
// Build: g++ -std=c++17 -O2 -shared -fPIC -o tests/emu/libtqb200_emu.so tests/emu/tqb_emu.cpp
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "../../torchquantum_b200/csrc/tqb_core.cuh"
#include "../../torchquantum_b200/csrc/tqb_block_build.cuh"
#include "../../torchquantum_b200/csrc/tqb_host_rng.h"
#include "../../include/tqb200.h"

using namespace tqb;

static char g_err[256] = "";
static int fail_msg(const char* m) { snprintf(g_err, sizeof(g_err), "%s", m); return -1; }

struct EmuExec {
  int nthreads;
  template <class F> void par(F f) { for (int t = 0; t < nthreads; ++t) f(t); }
  template <class F> void par_nosync(F f) { for (int t = 0; t < nthreads; ++t) f(t); }
  void async_copy8(float2* dst, const float2* src) { *dst = *src; }
  void async_copy16(float2* dst, const float2* src) { dst[0] = src[0]; dst[1] = src[1]; }
  void async_commit() {}
  void async_wait_prev() {}
  void async_wait_all() {}
  void global_add(float* p, float v) { *p += v; }
};

template <int R, bool BWD>
static void emu_blocks(BlockCtx c, int n, const PassGeom& g, unsigned ext_hi) {
  std::vector<unsigned char> smem(g.lay.total + 16);
  unsigned char* base = smem.data();
  base += (16 - ((uintptr_t)base & 15)) & 15;
  c.s_psi = reinterpret_cast<float2*>(base + g.lay.psi);
  c.s_lam = reinterpret_cast<float2*>(base + g.lay.lam);
  c.s_msh = reinterpret_cast<float*>(base + g.lay.msh);
  c.s_mps = reinterpret_cast<float*>(base + g.lay.mps);
  c.s_grad = reinterpret_cast<float*>(base + g.lay.grad);
  c.s_part = reinterpret_cast<float*>(base + g.lay.part);
  c.s_prog = reinterpret_cast<int32_t*>(base + g.lay.prog);
  c.s_scat = reinterpret_cast<u64*>(base + g.lay.scat);
  c.s_olut = reinterpret_cast<u64*>(base + g.lay.olut);
  // a small "grid" of persistent blocks so the multi-item loop and the double buffering are exercised
  const long long grid = g.n_items < 3 ? g.n_items : 3;
  for (long long blk = 0; blk < grid; ++blk) {
    EmuExec ex{c.nthreads};
    BlockCtx cb = c;  // run_cta advances the staged-buffer pointers of its context
    run_cta<R, BWD>(ex, cb, n, g.ntiles_log2, ext_hi, blk, grid, g.n_items, g.nbuf, g.lay.buf_stride);
  }
}

static int run_pass(bool bwd, const float* psi_src, const float* lam_src, float* psi, float* lam, int n, long long bsz,
                    const int32_t* prog, const int* desc, const float* msh, const float* mps, long long mps_stride,
                    float* gr, long long gr_stride, unsigned ext_hi) {
  const int T = desc[0], R = desc[1], L = desc[2], sh_max = desc[4], ps_max = desc[5], g_max = desc[6];
  if (T < 1 || T > TQB_MAX_T || T > n || R < 1 || R > TQB_MAX_R || R > T || L < 0 || L > T)
    return fail_msg("tqb pass: bad tile geometry");
  if (prog[TQB_H_MAGIC] != TQB_MAGIC || prog[TQB_H_T] != T || prog[TQB_H_R] != R || prog[TQB_H_L] != L ||
      prog[TQB_H_NSEG] != desc[3] || prog[TQB_H_SH_MAX] != sh_max || prog[TQB_H_PS_MAX] != ps_max ||
      prog[TQB_H_G_MAX] != g_max)
    return fail_msg("tqb pass: program header mismatch");
  if (sh_max > TQB_SEG_MAX_SH || ps_max > TQB_SEG_MAX_PS || g_max > TQB_SEG_MAX_G)
    return fail_msg("tqb pass: segment exceeds staging capacity");
  if (bsz <= 0) return 0;
  PassGeom g;
  pass_geometry(n, bsz, T, R, L, bwd, sh_max, ps_max, g_max, 200 * 1024, g);
  if (g.lay.total > 227 * 1024) return fail_msg("tqb pass: tile does not fit shared memory");
  BlockCtx c;
  c.T = T; c.L = L; c.S = g.S; c.tps = g.tps; c.nthreads = g.S * g.tps; c.bsz = (int)bsz;
  c.multi_tile = g.ntiles_log2 > 0;
  c.sample0 = 0; c.tile_ext = 0; c.outer_off = 0;
  c.sample_stride = 1ull << n;
  c.prog = prog; c.psi = reinterpret_cast<float2*>(psi); c.lam = reinterpret_cast<float2*>(lam);
  c.psi_src = reinterpret_cast<const float2*>(psi_src); c.lam_src = reinterpret_cast<const float2*>(lam_src);
  c.msh = msh; c.mps = mps; c.gr = gr; c.mps_stride = mps_stride; c.gr_stride = gr_stride;
  c.ps_cap = (ps_max + 3) & ~3; c.g_cap = (g_max + 3) & ~3;
  c.scale_sh = prog[TQB_H_SCALE_SH]; c.scale_ps = prog[TQB_H_SCALE_PS];
#define TQB_DISPATCH(RR) \
  case RR: if (bwd) emu_blocks<RR, true>(c, n, g, ext_hi); else emu_blocks<RR, false>(c, n, g, ext_hi); return 0;
  switch (R) { TQB_DISPATCH(1) TQB_DISPATCH(2) TQB_DISPATCH(3) TQB_DISPATCH(4) }
#undef TQB_DISPATCH
  return fail_msg("bad R");
}

typedef std::complex<float> cf;
typedef std::complex<double> cd;

extern "C" {

int tqb_abi_version(void) { return TQB_ABI_VERSION; }
const char* tqb_last_error(void) { return g_err; }

int tqb_pass_forward(const float* src, float* dst, int n, long long bsz, const int32_t* prog, const int* desc,
                     const float* msh, const float* mps, long long mps_stride, unsigned ext_hi, void*) {
  return run_pass(false, src, nullptr, dst, nullptr, n, bsz, prog, desc, msh, mps, mps_stride, nullptr, 0, ext_hi);
}
int tqb_pass_backward(const float* psi_src, const float* lam_src, float* psi, float* lam, int n, long long bsz,
                      const int32_t* prog, const int* desc, const float* msh, const float* mps, long long mps_stride,
                      float* gr, long long gr_stride, unsigned ext_hi, void*) {
  return run_pass(true, psi_src, lam_src, psi, lam, n, bsz, prog, desc, msh, mps, mps_stride, gr, gr_stride, ext_hi);
}

int tqb_build_mats(const int32_t* table, int nrec, int n_sh, long long bsz, const float* psh, const float* pps,
                   int ps_stride, const float* consts, float* msh, float* mps, long long mps_stride, void*) {
  for (int r = 0; r < n_sh; ++r) build_one(table + r * TQB_B_WORDS, 0, psh, pps, ps_stride, consts, msh, table);
  for (long long b = 0; b < bsz; ++b)
    for (int r = n_sh; r < nrec; ++r)
      build_one(table + r * TQB_B_WORDS, b, psh, pps, ps_stride, consts, mps + b * mps_stride, table);
  return 0;
}
int tqb_build_mats_bwd(const int32_t* table, int nrec, long long bsz, const float* psh, const float* pps,
                       int ps_stride, const float* gr, long long gr_stride, float* gsh, int sh_stride, float* gps,
                       void*) {
  for (long long b = 0; b < bsz; ++b)
    for (int r = 0; r < nrec; ++r)
      build_one_bwd(table + r * TQB_B_WORDS, b, psh, pps, ps_stride, gr + b * gr_stride,
                    gsh + b * sh_stride, gps + b * ps_stride);
  return 0;
}

int tqb_apply_dense(const float* in_, float* out_, int n, long long bsz, const float* mats_, long long mstride, int k,
                    const int* bitpos, int conj_t, void*) {
  const cf* in = reinterpret_cast<const cf*>(in_);
  cf* out = reinterpret_cast<cf*>(out_);
  const cf* mats = reinterpret_cast<const cf*>(mats_);
  const int d = 1 << k;
  for (long long b = 0; b < bsz; ++b)
    for (u64 amp = 0; amp < (1ull << n); ++amp) {
      u64 mask = 0; int row = 0;
      for (int q = 0; q < k; ++q) { mask |= 1ull << bitpos[q]; row = (row << 1) | (int)((amp >> bitpos[q]) & 1); }
      cf acc = 0;
      for (int col = 0; col < d; ++col) {
        u64 j = amp & ~mask;
        for (int q = 0; q < k; ++q) if ((col >> (k - 1 - q)) & 1) j |= 1ull << bitpos[q];
        cf m = conj_t ? std::conj(mats[b * mstride + (long long)col * d + row]) : mats[b * mstride + (long long)row * d + col];
        acc += m * in[(b << n) + j];
      }
      out[(b << n) + amp] = acc;
    }
  return 0;
}

int tqb_apply_dense_tc(const float* in_, float* out_, int n, long long bsz, const float* mat_, int k, const int* bitpos_host,
                       int conj_t, void* st) {
  if (k < 4 || k > 7 || n < 12) return fail_msg("tqb_apply_dense_tc: needs 4 <= k <= 7 and n_qubits >= 12");
  return tqb_apply_dense(in_, out_, n, bsz, mat_, 0, k, bitpos_host, conj_t, st);
}

int tqb_block_build(const int32_t* blocks, int n_blocks, const int32_t* ops, const float* consts, const float* psh, unsigned rank,
                    float* planes, void*) {
  for (int b = 0; b < n_blocks; ++b)
    for (int j = 0; j < 64; ++j) block_build_one(blocks, ops, consts, psh, rank, b, j, planes);
  return 0;
}

// host restatement of tqb_block6_kernel's arithmetic contract: out tile = (A_hi + A_lo) * tile, the block's bits and the
// index bits 0-5 exchanged on the way out when the block is transposing, the tile written contiguously (block bits ->
// index bits 6-11, outer bits compacted above) when it is rotating (fp32 data, double accumulation).  Same buffer chain
// as the CUDA library: rotating blocks alternate between dst and scratch, the last block leaves the state in dst.
int tqb_apply_blocks(const float* src_, float* dst_, float* scratch_, int n, long long bsz, const float* planes_all,
                     const int32_t* blocks_host, int n_blocks, int, void*) {
  if (n < 12) return fail_msg("block6 needs n_qubits >= 12");
  if (bsz <= 0 || n_blocks <= 0) return 0;
  const u64 N = 1ull << n;
  std::vector<cf> tmp(N);
  int flips = 0;
  for (int b = 1; b < n_blocks; ++b) flips += (blocks_host[(size_t)b * TQB_KB_WORDS + TQB_KB_FLAGS] & TQB_KBF_ROTATE) ? 1 : 0;
  const int rot0 = (blocks_host[TQB_KB_FLAGS] & TQB_KBF_ROTATE) ? 1 : 0;
  if ((flips + rot0) > 0 && scratch_ == nullptr && (flips > 0 || src_ == dst_))
    return fail_msg("tqb_apply_blocks: rotating blocks need a scratch state");
  float* bufs[2] = {dst_, scratch_};
  int at = flips & 1;
  const float* cur = src_;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int32_t* bh = blocks_host + (size_t)blk * TQB_KB_WORDS;
    const float* pl = planes_all + (size_t)blk * 2 * 128 * 64;
    int kpos[6];
    u64 kmask = 0;
    for (int k = 0; k < 6; ++k) {
      kpos[k] = bh[TQB_KB_KPOS + k];
      if (kpos[k] < 6 || kpos[k] >= n) return fail_msg("block bit out of range (must be >= 6)");
      if ((kmask >> kpos[k]) & 1) return fail_msg("block bits must be distinct");
      kmask |= 1ull << kpos[k];
    }
    const bool tr = bh[TQB_KB_FLAGS] & TQB_KBF_TRANSPOSE;
    const bool rotate = bh[TQB_KB_FLAGS] & TQB_KBF_ROTATE;
    if (blk > 0 && rotate) at ^= 1;
    float* outp = bufs[at];
    if (rotate && outp == cur) return fail_msg("tqb_apply_blocks: with an odd number of rotating blocks dst must differ from src (pass src as scratch)");
    std::vector<cd> U(64 * 64);
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j)
        U[i * 64 + j] = cd((double)pl[(2 * i) * 64 + j] + (double)pl[(128 + 2 * i) * 64 + j],
                           (double)pl[(2 * i + 1) * 64 + j] + (double)pl[(128 + 2 * i + 1) * 64 + j]);
    auto spread = [&](int pat) { u64 o = 0; for (int k = 0; k < 6; ++k) if ((pat >> k) & 1) o |= 1ull << kpos[k]; return o; };
    // destination of the tile's outer bits: compacted above bit 11 for a rotating block, unchanged otherwise
    auto outer_dst = [&](u64 base) {
      if (!rotate) return base;
      u64 o = 0; int r = 12;
      for (int bit = 6; bit < n; ++bit) if (!((kmask >> bit) & 1)) { o |= ((base >> bit) & 1ull) << r; ++r; }
      return o;
    };
    for (long long b = 0; b < bsz; ++b) {
      const cf* in = reinterpret_cast<const cf*>(cur) + (u64)b * N;
      cf* out = reinterpret_cast<cf*>(outp) + (u64)b * N;
      for (u64 base = 0; base < N; ++base) {
        if (base & (kmask | 63ull)) continue;
        const u64 ob = outer_dst(base);
        for (int nn = 0; nn < 64; ++nn) {
          cd v[64];
          for (int j = 0; j < 64; ++j) v[j] = cd(in[base | spread(j) | (u64)nn]);
          for (int i = 0; i < 64; ++i) {
            cd acc = 0;
            for (int j = 0; j < 64; ++j) acc += U[i * 64 + j] * v[j];
            const u64 blk_bits = rotate ? ((u64)(tr ? nn : i) << 6) : spread(tr ? nn : i);
            const u64 at_ = ob | blk_bits | (u64)(tr ? i : nn);
            tmp[at_] = cf((float)acc.real(), (float)acc.imag());
          }
        }
      }
      for (u64 a = 0; a < N; ++a) out[a] = tmp[a];
    }
    cur = outp;
  }
  return 0;
}

// host restatement of tqb_quadphase_kernel: the same wrapping uint32 turn arithmetic, sin / cos in double
int tqb_quadphase(float* state_, int n_local, long long bsz, const uint32_t* coef, int n_bits, unsigned rank, void*) {
  if (n_local < 12) return fail_msg("quadphase needs n_local >= 12");
  if (n_bits < n_local || n_bits > TQB_QP_BITS) return fail_msg("quadphase: n_bits must be in [n_local, TQB_QP_BITS]");
  cf* st = reinterpret_cast<cf*>(state_);
  const u64 N = 1ull << n_local;
  const uint32_t* alpha = coef + 1;
  const uint32_t* beta = coef + 1 + TQB_QP_BITS;
  for (u64 k = 0; k < N; ++k) {
    const u64 x = k | ((u64)rank << n_local);
    uint32_t t = coef[0];
    for (int b = 0; b < n_bits; ++b)
      if ((x >> b) & 1ull) {
        t += alpha[b];
        for (int a = 0; a < b; ++a)
          if ((x >> a) & 1ull) t += beta[a * TQB_QP_BITS + b];
      }
    const double ang = (double)(int32_t)t * (3.14159265358979323846 / 2147483648.0);
    const cd ph(std::cos(ang), std::sin(ang));
    for (long long b = 0; b < bsz; ++b) {
      const cd v = cd(st[(u64)b * N + k]) * ph;
      st[(u64)b * N + k] = cf((float)v.real(), (float)v.imag());
    }
  }
  return 0;
}

int tqb_dense_grad(const float* psi_, const float* lam_, int n, long long bsz, int k, const int* bitpos, float* G_, void*) {
  const cf* psi = reinterpret_cast<const cf*>(psi_);
  const cf* lam = reinterpret_cast<const cf*>(lam_);
  cf* G = reinterpret_cast<cf*>(G_);
  const int d = 1 << k;
  for (long long i = 0; i < bsz * d * d; ++i) G[i] = 0;
  for (long long b = 0; b < bsz; ++b)
    for (u64 amp = 0; amp < (1ull << n); ++amp) {
      u64 mask = 0; int row = 0;
      for (int q = 0; q < k; ++q) { mask |= 1ull << bitpos[q]; row = (row << 1) | (int)((amp >> bitpos[q]) & 1); }
      for (int col = 0; col < d; ++col) {
        u64 j = amp & ~mask;
        for (int q = 0; q < k; ++q) if ((col >> (k - 1 - q)) & 1) j |= 1ull << bitpos[q];
        G[(b * d + row) * d + col] += lam[(b << n) + amp] * std::conj(psi[(b << n) + j]);
      }
    }
  return 0;
}

static cf pauli_apply(const cf* psi, u64 k, u64 x, u64 z, int ny) {  // (P psi)[k]
  const u64 j = k ^ x;
  cf v = psi[j];
  static const cf iph[4] = {cf(1, 0), cf(0, 1), cf(-1, 0), cf(0, -1)};
  v *= iph[ny & 3];
  if (__builtin_popcountll(j & z) & 1) v = -v;
  return v;
}

int tqb_expval_pauli(const float* state, int n, long long bsz, const unsigned long long* xm, const unsigned long long* zm,
                     const int* ny, int nt, float* out, void*) {
  const cf* s = reinterpret_cast<const cf*>(state);
  for (long long b = 0; b < bsz; ++b)
    for (int t = 0; t < nt; ++t) {
      double acc = 0;
      for (u64 k = 0; k < (1ull << n); ++k) acc += (std::conj(s[(b << n) + k]) * pauli_apply(s + (b << n), k, xm[t], zm[t], ny[t])).real();
      if (n > 12) out[b * nt + t] += (float)acc; else out[b * nt + t] = (float)acc;  // same contract as the CUDA build
    }
  return 0;
}
int tqb_expval_pauli_bwd(const float* state, int n, long long bsz, const unsigned long long* xm,
                         const unsigned long long* zm, const int* ny, int nt, const float* gout, float* grad, void*) {
  const cf* s = reinterpret_cast<const cf*>(state);
  cf* g = reinterpret_cast<cf*>(grad);
  for (long long b = 0; b < bsz; ++b)
    for (u64 k = 0; k < (1ull << n); ++k) {
      cf acc = 0;
      for (int t = 0; t < nt; ++t) acc += 2.f * gout[b * nt + t] * pauli_apply(s + (b << n), k, xm[t], zm[t], ny[t]);
      g[(b << n) + k] = acc;
    }
  return 0;
}
int tqb_expval_z(const float* state, int n, long long bsz, float* out, void*) {
  const cf* s = reinterpret_cast<const cf*>(state);
  for (long long b = 0; b < bsz; ++b) {
    std::vector<double> acc(n + 1, 0.0);
    for (u64 k = 0; k < (1ull << n); ++k) {
      const double p = std::norm(cd(s[(b << n) + k]));
      acc[n] += p;
      for (int j = 0; j < n; ++j) acc[j] += ((k >> j) & 1) ? -p : p;
    }
    for (int j = 0; j <= n; ++j) { if (n > 12) out[b * (n + 1) + j] += (float)acc[j]; else out[b * (n + 1) + j] = (float)acc[j]; }
  }
  return 0;
}
int tqb_expval_z_bwd(const float* state, int n, long long bsz, const float* gout, float* grad, void*) {
  const cf* s = reinterpret_cast<const cf*>(state);
  cf* g = reinterpret_cast<cf*>(grad);
  for (long long b = 0; b < bsz; ++b)
    for (u64 k = 0; k < (1ull << n); ++k) {
      float w = gout[b * (n + 1) + n];
      for (int j = 0; j < n; ++j) w += ((k >> j) & 1) ? -gout[b * (n + 1) + j] : gout[b * (n + 1) + j];
      g[(b << n) + k] = 2.f * w * s[(b << n) + k];
    }
  return 0;
}
int tqb_probs(const float* state, int n, long long bsz, float* probs, void*) {
  for (long long i = 0; i < (bsz << n); ++i) {
    const double re = state[2 * i], im = state[2 * i + 1];
    const float m = (float)std::sqrt(re * re + im * im);
    probs[i] = m * m;
  }
  return 0;
}
int tqb_sample_counts(const float* probs, int n, long long bsz, const float* u, int shots, float* cdf, int* counts, void*) {
  const long long N = 1ll << n;
  for (long long b = 0; b < bsz; ++b) {
    float run = 0.f;
    for (long long k = 0; k < N; ++k) { run = k == 0 ? probs[b * N] : run + probs[b * N + k]; cdf[b * N + k] = run; }
    for (int s = 0; s < shots; ++s) {
      const float r = u[b * shots + s] * cdf[b * N + N - 1];
      long long lo = 0, hi = N - 1;
      while (lo < hi) { const long long mid = (lo + hi) >> 1; if (r < cdf[b * N + mid]) hi = mid; else lo = mid + 1; }
      counts[b * N + lo] += 1;
    }
  }
  return 0;
}
static inline float emu_tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
int tqb_gemm_prep(const float* x_, long long rows, long long cols, int mode, float* hi, float* lo, void*) {
  if (rows <= 0 || cols <= 0 || rows % 32 || cols % 32) return fail_msg("tqb_gemm_prep: rows and cols must be positive multiples of 32");
  if (mode != 1 && mode != 2) return fail_msg("tqb_gemm_prep: mode must be 1 (A of x^H) or 2 (Bt of x)");
  auto rna = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xffffe000u; float r; std::memcpy(&r, &u, 4); return r; };
  auto put = [&](long long row, long long col, float v) {   // planes are [.][2 rows] floats
    const float h = rna(v);
    hi[row * 2 * rows + col] = h;
    lo[row * 2 * rows + col] = rna(v - h);
  };
  for (long long r = 0; r < rows; ++r)
    for (long long c = 0; c < cols; ++c) {
      const float re = x_[2 * (r * cols + c)], im = x_[2 * (r * cols + c) + 1];
      if (mode == 1) {
        put(c, 2 * r, re); put(c, 2 * r + 1, -im);
      } else {
        put(2 * c, 2 * r, re); put(2 * c, 2 * r + 1, -im);
        put(2 * c + 1, 2 * r, im); put(2 * c + 1, 2 * r + 1, re);
      }
    }
  return 0;
}

int tqb_split_tf32(const float* x, float* hi, float* lo, long long n, void*) {
  for (long long i = 0; i < n; ++i) { hi[i] = emu_tf32_rna(x[i]); lo[i] = emu_tf32_rna(x[i] - hi[i]); }
  return 0;
}
long long tqb_gemm_workspace_floats(long long, long long, long long) { return 0; }
int tqb_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* bt_hi, const float* bt_lo, float* c, float*, long long M,
                    long long N, long long K, void*) {
  if (M % 128 || N % 256 || K % 32) return fail_msg("gemm: M % 128, N % 256, K % 32 must be 0");
  for (long long m = 0; m < M; ++m)
    for (long long n = 0; n < N; ++n) {
      float acc = 0.f;   // same order of the three products as the device kernel
      for (long long k = 0; k < K; ++k) acc += a_lo[m * K + k] * bt_hi[n * K + k];
      for (long long k = 0; k < K; ++k) acc += a_hi[m * K + k] * bt_lo[n * K + k];
      for (long long k = 0; k < K; ++k) acc += a_hi[m * K + k] * bt_hi[n * K + k];
      c[m * N + n] = acc;
    }
  return 0;
}
int tqb_exchange_p2p(const float*, float* const*, int, long long, int, const int*, int, void*) {
  return fail_msg("tqb_exchange_p2p: peer memory exists only on the GPU build (the gloo tests exchange through all_to_all)");
}
int tqb_host_uniforms(uint32_t* mt_state625_host, long long n, float* out_host) {
  if (n < 0 || !mt_state625_host || mt_state625_host[624] > 624u) return fail_msg("tqb_host_uniforms: bad arguments");
  tqb_mt19937_uniforms(mt_state625_host, n, out_host);
  return 0;
}
}  // extern "C"
