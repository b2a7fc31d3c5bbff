// sm_100a kernels + the C ABI of libtqb200.so (declared in include/tqb200.h).
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tqb_core.cuh"
#include "tqb_dense_tc.cuh"
#include "tqb_block_umma.cuh"
#include "tqb_block_build.cuh"
#include "tqb_gemm_umma.cuh"
#include "tqb_quadphase.cuh"
#include "tqb_host_rng.h"
#include "../../include/tqb200.h"

using namespace tqb;

static thread_local char g_err[256] = "";
static int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}
static int fail_msg(const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return -1;
}
#define TQB_CHECK_LAUNCH(name)                    \
  do {                                            \
    cudaError_t e_ = cudaGetLastError();          \
    if (e_ != cudaSuccess) return fail(name, e_); \
  } while (0)

// ---------------------------------------------------------------------------------------
// fused gate pass (forward / adjoint backward)
// ---------------------------------------------------------------------------------------
struct CudaExec {
  template <class F>
  __device__ __forceinline__ void par(F f) {
    f((int)threadIdx.x);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void par_nosync(F f) { f((int)threadIdx.x); }
  // global -> shared without a register round trip (LDGSTS); completion via commit/wait groups
  __device__ __forceinline__ void async_copy8(float2* dst, const float2* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
  }
  __device__ __forceinline__ void async_copy16(float2* dst, const float2* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
  }
  __device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
  __device__ __forceinline__ void async_wait_prev() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
  __device__ __forceinline__ void async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
  __device__ __forceinline__ void global_add(float* p, float v) { atomicAdd(p, v); }
};

// min CTAs/SM of the forward kernel: 2 (<= 128 registers, no spills) beat 3 (80 registers, spills) on every pass size
#ifndef TQB_FWD_MINB
#define TQB_FWD_MINB 2
#endif
// backward: 1 CTA/SM (255 registers, no spills) beat 2 (128 registers, 3.7 KB of spills): profiles/r01_kernel_variants.txt
#ifndef TQB_BWD_MINB
#define TQB_BWD_MINB 1
#endif
static int g_num_sms = 0;

struct PassArgs {
  int T, L, n, S, tps, bsz, nbuf, ps_cap, g_cap;
  unsigned ntiles_log2, ext_hi;
  long long n_items;
  const int32_t* prog;
  const float2* psi_src;
  const float2* lam_src;
  float2* psi;
  float2* lam;
  const float* msh;
  const float* mps;
  float* gr;
  long long mps_stride, gr_stride;
  unsigned o_psi, o_lam, o_msh, o_mps, o_grad, o_part, o_prog, o_scat, o_olut, buf_stride;
};

// SMALL: CTAs of <= 128 threads (states of <= 11 qubits held whole in shared memory, several samples per CTA).
// Backward (config-2 adjoint pass, profiles/r01_v3_notes.txt): with the per-type dispatcher 3 CTAs/SM (168 registers,
// 84 B of spills) = 2.28 ms beats 2 CTAs/SM (225 registers, no spills) = 2.47 ms; 4 CTAs/SM (128 registers) spills 200 B.
#ifndef TQB_BWD_MINB_SMALL
#define TQB_BWD_MINB_SMALL 3
#endif
// small-CTA forward: 6 CTAs/SM (80 registers, 56 B of spills) beat 4 (128 registers): config-2 forward 0.97 vs 1.03 ms,
// forward-only program 0.82 vs 0.91 ms; the 256-thread forward kernel is the other way round (3 CTAs/SM: 3.31 vs 3.00 ms)
#ifndef TQB_FWD_MINB_SMALL
#define TQB_FWD_MINB_SMALL 6
#endif
template <int R, bool BWD, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 128 : 256, SMALL ? (BWD ? TQB_BWD_MINB_SMALL : TQB_FWD_MINB_SMALL) : (BWD ? TQB_BWD_MINB : TQB_FWD_MINB))
tqb_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  BlockCtx c;
  c.T = a.T; c.L = a.L; c.S = a.S; c.tps = a.tps; c.nthreads = a.S * a.tps;
  c.bsz = a.bsz;
  c.multi_tile = a.ntiles_log2 > 0;
  c.sample0 = 0; c.tile_ext = 0; c.outer_off = 0;
  c.sample_stride = 1ull << a.n;
  c.prog = a.prog; c.psi_src = a.psi_src; c.lam_src = a.lam_src; c.psi = a.psi; c.lam = a.lam;
  c.msh = a.msh; c.mps = a.mps; c.gr = a.gr; c.mps_stride = a.mps_stride; c.gr_stride = a.gr_stride;
  c.ps_cap = a.ps_cap; c.g_cap = a.g_cap;
  c.scale_sh = a.prog[TQB_H_SCALE_SH]; c.scale_ps = a.prog[TQB_H_SCALE_PS];
  c.s_psi = reinterpret_cast<float2*>(smem + a.o_psi);
  c.s_lam = reinterpret_cast<float2*>(smem + a.o_lam);
  c.s_msh = reinterpret_cast<float*>(smem + a.o_msh);
  c.s_mps = reinterpret_cast<float*>(smem + a.o_mps);
  c.s_grad = reinterpret_cast<float*>(smem + a.o_grad);
  c.s_part = reinterpret_cast<float*>(smem + a.o_part);
  c.s_prog = reinterpret_cast<int32_t*>(smem + a.o_prog);
  c.s_scat = reinterpret_cast<u64*>(smem + a.o_scat);
  c.s_olut = reinterpret_cast<u64*>(smem + a.o_olut);
  CudaExec ex;
  run_cta<R, BWD>(ex, c, a.n, a.ntiles_log2, a.ext_hi, (long long)blockIdx.x, (long long)gridDim.x, a.n_items, a.nbuf,
                  (size_t)a.buf_stride);
}

template <int R, bool BWD, bool SMALL>
static int launch_pass_t(const PassArgs& a, size_t smem, cudaStream_t st) {
  // per instantiation AND per device (kernel attributes are per device; one process normally drives one GPU)
  static int cached_dev = -1, configured_smem = -1;
  static int occ_threads = 0, occ_smem = -1, occ_blocks = 0;
  const int threads = a.S * a.tps;
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (dev_now != cached_dev) { cached_dev = dev_now; configured_smem = -1; occ_threads = 0; occ_smem = -1; }
  if ((int)smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(tqb_pass_kernel<R, BWD, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(pass)", e);
    configured_smem = (int)smem;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  if (occ_threads != threads || occ_smem != (int)smem) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tqb_pass_kernel<R, BWD, SMALL>, threads, smem);
    if (e != cudaSuccess) return fail("cudaOccupancyMaxActiveBlocksPerMultiprocessor(pass)", e);
    occ_threads = threads; occ_smem = (int)smem; occ_blocks = nb < 1 ? 1 : nb;
  }
  long long grid = (long long)g_num_sms * occ_blocks;
  if (grid > a.n_items) grid = a.n_items;
  tqb_pass_kernel<R, BWD, SMALL><<<(unsigned)grid, threads, smem, st>>>(a);
  TQB_CHECK_LAUNCH("tqb_pass_kernel");
  return 0;
}

template <int R, bool BWD>
static int launch_pass(const PassArgs& a, size_t smem, cudaStream_t st) {
  return a.S * a.tps <= 128 ? launch_pass_t<R, BWD, true>(a, smem, st) : launch_pass_t<R, BWD, false>(a, smem, st);
}

static int run_pass(bool bwd, const float* psi_src, const float* lam_src, float* psi, float* lam, int n, long long bsz,
                    const int32_t* prog, const int* desc, const float* msh, const float* mps, long long mps_stride,
                    float* gr, long long gr_stride, unsigned ext_hi, void* stream) {
  const int T = desc[0], R = desc[1], L = desc[2], sh_max = desc[4], ps_max = desc[5], g_max = desc[6];
  if (T < 1 || T > TQB_MAX_T || T > n || R < 1 || R > TQB_MAX_R || R > T || L < 0 || L > T)
    return fail_msg("tqb pass: bad tile geometry");
  if (sh_max > TQB_SEG_MAX_SH || ps_max > TQB_SEG_MAX_PS || g_max > TQB_SEG_MAX_G)
    return fail_msg("tqb pass: segment exceeds staging capacity");
  if (bsz <= 0) return 0;
  if ((((uintptr_t)psi_src | (uintptr_t)psi | (uintptr_t)lam_src | (uintptr_t)lam) & 15) != 0)
    return fail_msg("tqb pass: state pointers must be 16-byte aligned");
  PassGeom g;
  pass_geometry(n, bsz, T, R, L, bwd, sh_max, ps_max, g_max, 200 * 1024, g);
  if (g.lay.total > 227 * 1024) return fail_msg("tqb pass: tile does not fit shared memory");
  PassArgs a;
  a.T = T; a.L = L; a.n = n; a.S = g.S; a.tps = g.tps; a.bsz = (int)bsz; a.nbuf = g.nbuf;
  a.ps_cap = (ps_max + 3) & ~3; a.g_cap = (g_max + 3) & ~3;
  a.ntiles_log2 = g.ntiles_log2; a.ext_hi = ext_hi; a.n_items = g.n_items;
  a.prog = prog; a.psi = reinterpret_cast<float2*>(psi); a.lam = reinterpret_cast<float2*>(lam);
  a.psi_src = reinterpret_cast<const float2*>(psi_src); a.lam_src = reinterpret_cast<const float2*>(lam_src);
  a.msh = msh; a.mps = mps; a.gr = gr; a.mps_stride = mps_stride; a.gr_stride = gr_stride;
  a.o_psi = (unsigned)g.lay.psi; a.o_lam = (unsigned)g.lay.lam; a.o_msh = (unsigned)g.lay.msh; a.o_mps = (unsigned)g.lay.mps;
  a.o_grad = (unsigned)g.lay.grad; a.o_part = (unsigned)g.lay.part; a.o_prog = (unsigned)g.lay.prog; a.o_scat = (unsigned)g.lay.scat; a.o_olut = (unsigned)g.lay.olut;
  a.buf_stride = (unsigned)g.lay.buf_stride;
  cudaStream_t st = (cudaStream_t)stream;
#define TQB_DISPATCH(RR)                                             \
  case RR:                                                           \
    return bwd ? launch_pass<RR, true>(a, g.lay.total, st) : launch_pass<RR, false>(a, g.lay.total, st);
  switch (R) {
    TQB_DISPATCH(1)
    TQB_DISPATCH(2)
    TQB_DISPATCH(3)
    TQB_DISPATCH(4)
  }
#undef TQB_DISPATCH
  return fail_msg("tqb pass: bad R");
}

// ---------------------------------------------------------------------------------------
// parameter -> matrix builder and its backward
// ---------------------------------------------------------------------------------------
// records [0, n_sh) write the shared row once; records [n_sh, nrec) write every sample's row
__global__ void tqb_build_kernel(const int32_t* table, int nrec, int n_sh, long long bsz, const float* params_sh,
                                 const float* params_ps, int ps_stride, const float* consts, float* msh, float* mps,
                                 long long mps_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_ps = nrec - n_sh;
  if (i < n_sh) {
    build_one(table + i * TQB_B_WORDS, 0, params_sh, params_ps, ps_stride, consts, msh, table);
    return;
  }
  const long long k = i - n_sh;
  if (k >= bsz * n_ps) return;
  const int r = n_sh + (int)(k % n_ps);
  const long long b = k / n_ps;
  build_one(table + r * TQB_B_WORDS, b, params_sh, params_ps, ps_stride, consts, mps + b * mps_stride, table);
}

__global__ void tqb_build_bwd_kernel(const int32_t* table, int nrec, long long bsz, const float* params_sh,
                                     const float* params_ps, int ps_stride, const float* gr, long long gr_stride,
                                     float* gsh, int sh_stride, float* gps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bsz * nrec) return;
  const int r = (int)(i % nrec);
  const long long b = i / nrec;
  build_one_bwd(table + r * TQB_B_WORDS, b, params_sh, params_ps, ps_stride, gr + b * gr_stride,
                gsh + b * sh_stride, gps + b * ps_stride);
}

// ---------------------------------------------------------------------------------------
// generic dense k-qubit gate, out of place (rare: qft / qubitunitary on >= 3 wires)
// ---------------------------------------------------------------------------------------
// one CTA per block, one thread per column of the block unitary (tqb_block_build.cuh)
__global__ void __launch_bounds__(64) tqb_block_build_kernel(const int32_t* blocks, const int32_t* ops, const float* consts,
                                                             const float* params_sh, unsigned rank, float* planes) {
  block_build_one(blocks, ops, consts, params_sh, rank, (int)blockIdx.x, (int)threadIdx.x, planes);
}

__global__ void tqb_dense_kernel(const float2* in, float2* out, int n, long long total, const float2* mats,
                                 long long mat_stride, int k, const int* bitpos /* msb first */, int conj_t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const u64 amp = (u64)i & ((1ull << n) - 1ull);
  const long long b = i >> n;
  u64 mask = 0;
  int row = 0;
  for (int q = 0; q < k; ++q) {
    const int p = bitpos[q];
    mask |= 1ull << p;
    row = (row << 1) | (int)((amp >> p) & 1ull);
  }
  const u64 base = amp & ~mask;
  const int d = 1 << k;
  const float2* m = mats + b * mat_stride;
  const float2* src = in + (b << n);
  float2 acc = make_float2(0.f, 0.f);
  for (int col = 0; col < d; ++col) {
    u64 j = base;
    for (int q = 0; q < k; ++q)
      if ((col >> (k - 1 - q)) & 1) j |= 1ull << bitpos[q];
    const float2 v = src[j];
    float2 mv = conj_t ? m[(long long)col * d + row] : m[(long long)row * d + col];
    if (conj_t) mv.y = -mv.y;
    acc = cadd(acc, cmul(mv, v));
  }
  out[i] = acc;
}

// G[b][r][c] = sum_rest lam[b][r,rest] conj(psi[b][c,rest])  (one block per (b, r, c))
__global__ void tqb_dense_grad_kernel(const float2* psi, const float2* lam, int n, int k, const int* bitpos,
                                      float2* G) {
  const int d = 1 << k;
  const int c = blockIdx.x % d, r = (blockIdx.x / d) % d;
  const long long b = blockIdx.x / (d * d);
  u64 rbits = 0, cbits = 0;
  u32 sp[32];
  for (int q = 0; q < k; ++q) {
    if ((r >> (k - 1 - q)) & 1) rbits |= 1ull << bitpos[q];
    if ((c >> (k - 1 - q)) & 1) cbits |= 1ull << bitpos[q];
    sp[q] = bitpos[q];
  }
  for (int a2 = 0; a2 < k; ++a2)
    for (int b2 = 0; b2 + 1 < k - a2; ++b2)
      if (sp[b2] > sp[b2 + 1]) { u32 t = sp[b2]; sp[b2] = sp[b2 + 1]; sp[b2 + 1] = t; }
  const u64 nrest = 1ull << (n - k);
  float2 acc = make_float2(0.f, 0.f);
  for (u64 w = threadIdx.x; w < nrest; w += blockDim.x) {
    u64 base = w;
    for (int q = 0; q < k; ++q) base = ((base >> sp[q]) << (sp[q] + 1)) | (base & ((1ull << sp[q]) - 1ull));
    const float2 l = lam[(b << n) + (base | rbits)], p = psi[(b << n) + (base | cbits)];
    acc = cadd(acc, cmulc(p, l));
  }
  __shared__ float2 red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] = cadd(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) G[blockIdx.x] = red[0];
}

// ---------------------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define TQB_TERM_BATCH 8
// out[b][t] (+)= <psi_b| P_t |psi_b>,  P|k> = i^ny (-1)^popc(k&z) |k^x>
// TB terms share every load of psi[k]; TB = 1 serves the reference API (one call per Pauli string)
template <int TB>
__global__ void __launch_bounds__(256) tqb_expval_pauli_kernel(const float2* state, int n, const u64* xm,
                                                                const u64* zm, const int* ny, int nterms,
                                                                float* out, int chunk_log2, int use_atomic) {
  const long long b = blockIdx.y;
  const float2* psi = state + (b << n);
  const u64 k0 = (u64)blockIdx.x << chunk_log2;
  const u64 k1 = k0 + (1ull << chunk_log2);
  __shared__ float red[8][TB];
  for (int t0 = 0; t0 < nterms; t0 += TB) {
    float acc[TB];
    u64 x[TB], z[TB];
    int ph[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      acc[t] = 0.f;
      const int tt = t0 + t < nterms ? t0 + t : nterms - 1;
      x[t] = xm[tt]; z[t] = zm[tt]; ph[t] = ny[tt] & 3;
    }
    for (u64 k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      const float2 a = psi[k];
#pragma unroll
      for (int t = 0; t < TB; ++t) {
        const float2 c = x[t] ? psi[k ^ x[t]] : a;
        float2 w = cmulc(c, a);  // conj(psi[k^x]) psi[k]
        float v = (ph[t] == 0) ? w.x : (ph[t] == 1) ? -w.y : (ph[t] == 2) ? -w.x : w.y;
        if (__popcll(k & z[t]) & 1) v = -v;
        acc[t] += v;
      }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      const float s = warp_sum(acc[t]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = s;
    }
    __syncthreads();
    if (threadIdx.x < TB && t0 + (int)threadIdx.x < nterms) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
      float* dst = out + b * nterms + t0 + threadIdx.x;
      if (use_atomic) atomicAdd(dst, s); else *dst = s;
    }
    __syncthreads();
  }
}

// Small states (n <= 13: a sample is at most 64 KiB): one WARP per sample, 128-bit loads of amplitude pairs, no
// block-level synchronisation.  The batched configs (cfg 1/2: thousands of 4-10 qubit samples) spend their read-out
// here; the block-per-chunk kernel above leaves most lanes idle on them.
template <int TB>
__global__ void __launch_bounds__(256) tqb_expval_pauli_warp_kernel(const float2* state, int n, const u64* xm,
                                                                     const u64* zm, const int* ny, int nterms,
                                                                     float* out, long long bsz) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (b >= bsz) return;
  const float4* psi4 = reinterpret_cast<const float4*>(state + (b << n));
  const u32 npairs = 1u << (n - 1);
  for (int t0 = 0; t0 < nterms; t0 += TB) {
    float acc[TB];
    u32 x[TB], z[TB];
    int ph[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      acc[t] = 0.f;
      const int tt = t0 + t < nterms ? t0 + t : nterms - 1;
      x[t] = (u32)xm[tt]; z[t] = (u32)zm[tt]; ph[t] = ny[tt] & 3;
    }
#pragma unroll 4
    for (u32 p = lane; p < npairs; p += 32u) {
      const float4 a = psi4[p];
      const float m0 = fmaf(a.y, a.y, a.x * a.x), m1 = fmaf(a.w, a.w, a.z * a.z);
#pragma unroll
      for (int t = 0; t < TB; ++t) {
        // (the branches are uniform over the warp: every lane works on the same term)
        if (x[t] == 0u) {            // Z / I string: +- |psi_k|^2, no partner amplitude, no complex product
          const u32 par = (u32)__popc((2u * p) & z[t]);
          acc[t] += ((par & 1u) ? -m0 : m0) + (((par + (z[t] & 1u)) & 1u) ? -m1 : m1);
          continue;
        }
        if (z[t] == 0u && ph[t] == 0) {   // X / I string: Re(conj(psi[k^x]) psi[k]), no sign
          float4 c = psi4[p ^ (x[t] >> 1)];
          if (x[t] & 1u) c = make_float4(c.z, c.w, c.x, c.y);
          acc[t] += fmaf(c.y, a.y, c.x * a.x) + fmaf(c.w, a.w, c.z * a.z);
          continue;
        }
        float4 c = psi4[p ^ (x[t] >> 1)];
        if (x[t] & 1u) c = make_float4(c.z, c.w, c.x, c.y);
        // w = conj(psi[k^x]) psi[k] for k = 2p, 2p+1
        const float w0x = fmaf(c.y, a.y, c.x * a.x), w0y = fmaf(-c.y, a.x, c.x * a.y);
        const float w1x = fmaf(c.w, a.w, c.z * a.z), w1y = fmaf(-c.w, a.z, c.z * a.w);
        float v0 = (ph[t] == 0) ? w0x : (ph[t] == 1) ? -w0y : (ph[t] == 2) ? -w0x : w0y;
        float v1 = (ph[t] == 0) ? w1x : (ph[t] == 1) ? -w1y : (ph[t] == 2) ? -w1x : w1y;
        const u32 par = (u32)__popc((2u * p) & z[t]);
        if (par & 1u) v0 = -v0;
        if ((par + (z[t] & 1u)) & 1u) v1 = -v1;
        acc[t] += v0 + v1;
      }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      const float s = warp_sum(acc[t]);
      if (lane == 0 && t0 + t < nterms) out[b * nterms + t0 + t] = s;
    }
  }
}

// grad[b][k] = sum_t 2 g[b][t] (P_t psi)[k],  (P psi)[k] = i^ny (-1)^popc((k^x)&z) psi[k^x]
__global__ void tqb_expval_pauli_bwd_kernel(const float2* state, int n, long long total, const u64* xm,
                                            const u64* zm, const int* ny, int nterms, const float* gout,
                                            float2* grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const u64 k = (u64)i & ((1ull << n) - 1ull);
  const long long b = i >> n;
  const float2* psi = state + (b << n);
  float2 acc = make_float2(0.f, 0.f);
  for (int t = 0; t < nterms; ++t) {
    const u64 j = k ^ xm[t];
    float2 v = psi[j];
    const int ph = ny[t] & 3;
    if (ph == 1) v = make_float2(-v.y, v.x);
    else if (ph == 2) v = make_float2(-v.x, -v.y);
    else if (ph == 3) v = make_float2(v.y, -v.x);
    float g = 2.f * gout[b * nterms + t];
    if (__popcll(j & zm[t]) & 1) g = -g;
    acc.x += g * v.x; acc.y += g * v.y;
  }
  grad[i] = acc;
}

// Same for small states (n <= 13): one warp per sample, amplitude pairs as 128-bit loads / stores, the term table
// (masks, phase, 2 g[b][t]) staged in shared memory 32 terms at a time.
__global__ void __launch_bounds__(256) tqb_expval_pauli_bwd_warp_kernel(const float2* state, int n, long long bsz,
                                                                         const u64* xm, const u64* zm, const int* ny,
                                                                         int nterms, const float* gout, float2* grad) {
  __shared__ u32 s_x[8][32], s_z[8][32], s_ph[8][32];
  __shared__ float s_g[8][32];
  const int wib = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  const long long b = (long long)blockIdx.x * 8 + wib;
  const bool live = b < bsz;
  const float4* psi4 = reinterpret_cast<const float4*>(state + ((live ? b : 0) << n));
  float4* out4 = reinterpret_cast<float4*>(grad + ((live ? b : 0) << n));
  const u32 npairs = 1u << (n - 1);
  for (int t0 = 0; t0 < nterms; t0 += 32) {
    const int nt = nterms - t0 < 32 ? nterms - t0 : 32;
    __syncwarp();
    if ((int)lane < nt) {
      s_x[wib][lane] = (u32)xm[t0 + lane]; s_z[wib][lane] = (u32)zm[t0 + lane]; s_ph[wib][lane] = (u32)(ny[t0 + lane] & 3);
      s_g[wib][lane] = live ? 2.f * gout[b * nterms + t0 + lane] : 0.f;
    }
    __syncwarp();
    if (!live) continue;
    for (u32 p = lane; p < npairs; p += 32u) {
      float4 acc = t0 == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : out4[p];
      float d0 = 0.f, d1 = 0.f;                        // Z / I strings: real factors on psi[2p], psi[2p + 1]
      for (int t = 0; t < nt; ++t) {
        const u32 x = s_x[wib][t], z = s_z[wib][t], ph = s_ph[wib][t];
        const float g = s_g[wib][t];
        if (x == 0u) {               // (uniform over the warp) diagonal string: no partner amplitude
          const u32 par = (u32)__popc((2u * p) & z);
          d0 += (par & 1u) ? -g : g;
          d1 += ((par + (z & 1u)) & 1u) ? -g : g;
          continue;
        }
        float4 v = psi4[p ^ (x >> 1)];                 // psi[j], j = k ^ x for k = 2p, 2p + 1
        if (x & 1u) v = make_float4(v.z, v.w, v.x, v.y);
        // i^ny psi[j]
        if (ph == 1u) v = make_float4(-v.y, v.x, -v.w, v.z);
        else if (ph == 2u) v = make_float4(-v.x, -v.y, -v.z, -v.w);
        else if (ph == 3u) v = make_float4(v.y, -v.x, v.w, -v.z);
        // sign (-1)^popc(j & z) for j = (2p ^ x) and (2p + 1) ^ x
        const u32 j0 = (2u * p) ^ x;
        const u32 par0 = (u32)__popc(j0 & z) & 1u, par1 = (u32)__popc((j0 ^ 1u) & z) & 1u;
        const float g0 = par0 ? -g : g, g1 = par1 ? -g : g;
        acc.x = fmaf(g0, v.x, acc.x); acc.y = fmaf(g0, v.y, acc.y);
        acc.z = fmaf(g1, v.z, acc.z); acc.w = fmaf(g1, v.w, acc.w);
      }
      const float4 a = psi4[p];
      acc.x = fmaf(d0, a.x, acc.x); acc.y = fmaf(d0, a.y, acc.y);
      acc.z = fmaf(d1, a.z, acc.z); acc.w = fmaf(d1, a.w, acc.w);
      out4[p] = acc;
    }
  }
}

// out[b][j] (+)= sum_k |psi_k|^2 (1 - 2 bit_j(k)) for j < n ; out[b][n] (+)= sum_k |psi_k|^2
__global__ void __launch_bounds__(256) tqb_expval_z_kernel(const float2* state, int n, float* out, int chunk_log2,
                                                            int use_atomic) {
  const long long b = blockIdx.y;
  const float2* psi = state + (b << n);
  const u64 k0 = (u64)blockIdx.x << chunk_log2;
  const int nloc = chunk_log2;  // bits that vary inside this block's chunk
  float tot = 0.f;
  float s1[TQB_MAX_T + 2];      // sum over amplitudes with local bit j set (chunk_log2 <= 14)
#pragma unroll
  for (int j = 0; j < TQB_MAX_T + 2; ++j) s1[j] = 0.f;
  for (u32 k = threadIdx.x; k < (1u << nloc); k += blockDim.x) {
    const float2 a = psi[k0 + k];
    const float p = a.x * a.x + a.y * a.y;
    tot += p;
#pragma unroll
    for (int j = 0; j < TQB_MAX_T + 2; ++j) if (j < nloc && ((k >> j) & 1u)) s1[j] += p;
  }
  __shared__ float red[TQB_MAX_T + 3];
  if (threadIdx.x < TQB_MAX_T + 3) red[threadIdx.x] = 0.f;
  __syncthreads();
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[TQB_MAX_T + 2], tot);
#pragma unroll
  for (int j = 0; j < TQB_MAX_T + 2; ++j) {
    if (j < nloc) {
      const float s = warp_sum(s1[j]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], s);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x <= n) {
    const int j = threadIdx.x;
    const float total = red[TQB_MAX_T + 2];
    float v;
    if (j == n) v = total;
    else if (j < nloc) v = total - 2.f * red[j];
    else v = ((k0 >> j) & 1ull) ? -total : total;
    float* dst = out + b * (n + 1) + j;
    if (use_atomic) atomicAdd(dst, v); else *dst = v;
  }
}

// grad[b][k] = 2 psi[k] (g[b][n] + sum_j g[b][j] (1 - 2 bit_j(k)))
__global__ void tqb_expval_z_bwd_kernel(const float2* state, int n, long long total, const float* gout, float2* grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const u64 k = (u64)i & ((1ull << n) - 1ull);
  const long long b = i >> n;
  const float* g = gout + b * (n + 1);
  float w = g[n];
  for (int j = 0; j < n; ++j) w += ((k >> j) & 1ull) ? -g[j] : g[j];
  const float2 a = state[i];
  grad[i] = make_float2(2.f * w * a.x, 2.f * w * a.y);
}

// probabilities exactly as numpy does on the host: m = float32(|psi|) (correctly rounded hypot),
// p = float32(m*m)   (measurement/measurements.py:47-55)
__global__ void tqb_probs_kernel(const float2* state, long long total, float* probs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float2 a = state[i];
  const float m = (float)sqrt((double)a.x * (double)a.x + (double)a.y * (double)a.y);
  probs[i] = __fmul_rn(m, m);
}

// sequential float32 running sum per sample (same rounding order as itertools.accumulate)
__global__ void tqb_cdf_kernel(const float* probs, long long bsz, long long N, float* cdf) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bsz) return;
  const float* p = probs + b * N;
  float* c = cdf + b * N;
  float run = 0.f;
  for (long long k = 0; k < N; ++k) {
    run = k == 0 ? p[0] : __fadd_rn(run, p[k]);
    c[k] = run;
  }
}

// counts[b][bisect_right(cdf_b, float32(u)*total, 0, N-1)] += 1   (CPython random.choices)
__global__ void tqb_sample_kernel(const float* cdf, const float* u, long long bsz, long long N, int shots,
                                  int* counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bsz * shots) return;
  const long long b = i / shots;
  const float* c = cdf + b * N;
  const float r = __fmul_rn(u[i], c[N - 1]);
  long long lo = 0, hi = N - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (r < c[mid]) hi = mid; else lo = mid + 1;
  }
  atomicAdd(counts + b * N + lo, 1);
}

// x -> TF32 hi / lo parts (round to nearest on the bit pattern, as the block kernel's splitters do): operand planes of the
// 3xTF32 GEMM
__global__ void __launch_bounds__(256) tqb_split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo,
                                                             long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    h.x = __uint_as_float(tqb_umma::tf32_rna(v.x)); h.y = __uint_as_float(tqb_umma::tf32_rna(v.y));
    h.z = __uint_as_float(tqb_umma::tf32_rna(v.z)); h.w = __uint_as_float(tqb_umma::tf32_rna(v.w));
    l.x = __uint_as_float(tqb_umma::tf32_rna(v.x - h.x)); l.y = __uint_as_float(tqb_umma::tf32_rna(v.y - h.y));
    l.z = __uint_as_float(tqb_umma::tf32_rna(v.z - h.z)); l.w = __uint_as_float(tqb_umma::tf32_rna(v.w - h.w));
    hi[i] = h;
    lo[i] = l;
  }
}

// Operand preparation of the complex GEMM in ONE pass over the complex matrix x [rows][cols] (replaces conj / transpose /
// stack / split chains of 5-7 elementwise kernels): the real operand the tensor-core GEMM wants, already cut into its TF32
// hi / lo planes.  32 x 32 tiles go through shared memory (reads contiguous along cols, writes contiguous along rows).
//   MODE 1  A  of x^H:  out[c][2r + v]      = (re, -im)[v] of x[r][c]                      out: [cols][2 rows]
//   MODE 2  Bt of x:    out[2c][2r + v]     = (re, -im)[v] of x[r][c],
//                       out[2c + 1][2r + v] = (im,  re)[v] of x[r][c]                      out: [2 cols][2 rows]
template <int MODE>
__global__ void __launch_bounds__(256) tqb_gemm_prep_kernel(const float2* __restrict__ x, long long rows, long long cols,
                                                            float2* __restrict__ hi, float2* __restrict__ lo) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = x[(r0 + ty + 8 * i) * cols + c0 + tx];
  __syncthreads();
  auto split = [](float v, float& h, float& l) {
    h = __uint_as_float(tqb_umma::tf32_rna(v));
    l = __uint_as_float(tqb_umma::tf32_rna(v - h));
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = tile[tx][ty + 8 * i];        // x[r0 + tx][c0 + ty + 8 i]
    const long long c = c0 + ty + 8 * i, r = r0 + tx;
    float2 h, l;
    split(v.x, h.x, l.x);
    split(-v.y, h.y, l.y);
    if (MODE == 1) {
      hi[c * rows + r] = h;
      lo[c * rows + r] = l;
    } else {
      hi[(2 * c) * rows + r] = h;
      lo[(2 * c) * rows + r] = l;
      hi[(2 * c + 1) * rows + r] = make_float2(-h.y, h.x);     // (im, re): the split of -(-im) and of re
      lo[(2 * c + 1) * rows + r] = make_float2(-l.y, l.x);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K10: global <-> local qubit exchange over peer memory (NVLink P2P stores), pack and transfer in ONE kernel.
// Rank r swaps its g rank-index bits with the g local index bits `vpos` (any positions, ascending): amplitude i of the
// local slab goes to rank v = (bits of i at vpos) at local index i with those bits replaced by the bits of r.  Every
// thread moves one 16-byte vector (VEC = 2 amplitudes; VEC = 1 when bit 0 is a victim) straight into the destination
// slab of the peer; a warp writes 512 contiguous bytes to one peer whenever the lowest victim bit is >= 6, which is what
// the victim choice in sharded.py prefers.  No staging buffer, no separate pack pass, no NCCL.
// ---------------------------------------------------------------------------------------
struct ExchangeArgs {
  float* dst[16];        // destination slab of every rank (peer-mapped device pointers), dst[rank] is local
  int vpos[4];
  int g, rank, nl;
};

template <int VEC>
__global__ void __launch_bounds__(256) tqb_exchange_kernel(const float* __restrict__ src, const ExchangeArgs a, long long total_vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const unsigned long long slab_mask = (1ull << a.nl) - 1ull;
  // Iteration order: the destination rank must vary FAST (every 1 KiB chunk of a warp goes to the next rank), otherwise the
  // grid sweeps the slab region by region and first does all local copies, then all remote ones: the NVLink and the local
  // HBM phases would run one after the other.  When every victim bit is >= 6 the chunk counter's low g bits are therefore
  // routed to the victim positions and its other bits fill the remaining index bits in order (a bit deposit with <= 4 holes).
  const bool interleave = VEC == 2 && a.vpos[0] >= 6;
  // source vector index and destination (rank, vector index) of iteration `it`
  auto route = [&](long long it, unsigned long long& from, int& to, unsigned long long& out) {
    unsigned long long flat = (unsigned long long)it * VEC;   // amplitude index within [bsz][2^nl]
    if (interleave) {
      const unsigned long long c = (unsigned long long)it >> 5;       // 64-amplitude chunk counter
      unsigned long long x = c >> a.g;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < a.g) {
          const int pos = a.vpos[k] - 6;
          x = ((x >> pos) << (pos + 1)) | (x & ((1ull << pos) - 1ull)) | (((c >> k) & 1ull) << pos);
        }
      flat = (x << 6) | (((unsigned long long)it & 31ull) << 1);
    }
    unsigned long long i = flat & slab_mask;
    to = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < a.g) {
        to |= (int)((i >> a.vpos[k]) & 1ull) << k;
        i = (i & ~(1ull << a.vpos[k])) | ((unsigned long long)((a.rank >> k) & 1) << a.vpos[k]);
      }
    from = flat;
    out = (flat & ~slab_mask) | i;
  };
  // four vectors per thread and trip: all loads first, then the (peer) stores
  constexpr int U = 4;
  long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; it + (U - 1) * stride < total_vec; it += U * stride) {
    unsigned long long from[U], out[U];
    int to[U];
#pragma unroll
    for (int u = 0; u < U; ++u) route(it + u * stride, from[u], to[u], out[u]);
    if (VEC == 2) {
      float4 val[U];
#pragma unroll
      for (int u = 0; u < U; ++u) val[u] = __ldcs(reinterpret_cast<const float4*>(src) + (from[u] >> 1));
#pragma unroll
      for (int u = 0; u < U; ++u) reinterpret_cast<float4*>(a.dst[to[u]])[out[u] >> 1] = val[u];
    } else {
      float2 val[U];
#pragma unroll
      for (int u = 0; u < U; ++u) val[u] = __ldcs(reinterpret_cast<const float2*>(src) + from[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) reinterpret_cast<float2*>(a.dst[to[u]])[out[u]] = val[u];
    }
  }
  for (; it < total_vec; it += stride) {
    unsigned long long from, out;
    int to;
    route(it, from, to, out);
    if (VEC == 2) {
      reinterpret_cast<float4*>(a.dst[to])[out >> 1] = __ldcs(reinterpret_cast<const float4*>(src) + (from >> 1));
    } else {
      reinterpret_cast<float2*>(a.dst[to])[out] = __ldcs(reinterpret_cast<const float2*>(src) + from);
    }
  }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int tqb_abi_version(void) { return TQB_ABI_VERSION; }
const char* tqb_last_error(void) { return g_err; }

int tqb_pass_forward(const float* src, float* dst, int n_qubits, long long bsz, const int32_t* prog_dev,
                     const int* desc_host, const float* msh, const float* mps, long long mps_stride, unsigned ext_hi,
                     void* stream) {
  return run_pass(false, src, nullptr, dst, nullptr, n_qubits, bsz, prog_dev, desc_host, msh, mps, mps_stride, nullptr, 0,
                  ext_hi, stream);
}

int tqb_pass_backward(const float* psi_src, const float* lam_src, float* psi_dst, float* lam_dst, int n_qubits,
                      long long bsz, const int32_t* prog_dev, const int* desc_host, const float* msh, const float* mps,
                      long long mps_stride, float* gr, long long gr_stride, unsigned ext_hi, void* stream) {
  return run_pass(true, psi_src, lam_src, psi_dst, lam_dst, n_qubits, bsz, prog_dev, desc_host, msh, mps, mps_stride, gr,
                  gr_stride, ext_hi, stream);
}

int tqb_build_mats(const int32_t* table_dev, int n_records, int n_shared_records, long long bsz, const float* params_sh,
                   const float* params_ps, int ps_stride, const float* consts, float* msh, float* mps,
                   long long mps_stride, void* stream) {
  if (n_records <= 0 || bsz <= 0) return 0;
  const long long total = n_shared_records + bsz * (long long)(n_records - n_shared_records);
  tqb_build_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      table_dev, n_records, n_shared_records, bsz, params_sh, params_ps, ps_stride, consts, msh, mps, mps_stride);
  TQB_CHECK_LAUNCH("tqb_build_kernel");
  return 0;
}

int tqb_build_mats_bwd(const int32_t* table_dev, int n_records, long long bsz, const float* params_sh,
                       const float* params_ps, int ps_stride, const float* gr, long long gr_stride,
                       float* gparams_sh_partial, int sh_stride, float* gparams_ps, void* stream) {
  if (n_records <= 0 || bsz <= 0) return 0;
  const long long total = bsz * n_records;
  tqb_build_bwd_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      table_dev, n_records, bsz, params_sh, params_ps, ps_stride, gr, gr_stride, gparams_sh_partial, sh_stride,
      gparams_ps);
  TQB_CHECK_LAUNCH("tqb_build_bwd_kernel");
  return 0;
}

int tqb_apply_dense(const float* in, float* out, int n_qubits, long long bsz, const float* mats,
                    long long mat_stride_complex, int k, const int* bitpos_dev, int conj_transpose, void* stream) {
  const long long total = bsz << n_qubits;
  if (total <= 0) return 0;
  tqb_dense_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), n_qubits, total,
      reinterpret_cast<const float2*>(mats), mat_stride_complex, k, bitpos_dev, conj_transpose);
  TQB_CHECK_LAUNCH("tqb_dense_kernel");
  return 0;
}

int tqb_apply_dense_tc(const float* in, float* out, int n_qubits, long long bsz, const float* mat, int k,
                       const int* bitpos_host, int conj_transpose, void* stream) {
  if (k < 4 || k > 7 || n_qubits < TQB_TC_T) return fail_msg("tqb_apply_dense_tc: needs 4 <= k <= 7 and n_qubits >= 12");
  if (bsz <= 0) return 0;
  DenseTcArgs a;
  a.in = reinterpret_cast<const float2*>(in); a.out = reinterpret_cast<float2*>(out);
  a.mat = reinterpret_cast<const float2*>(mat);
  a.n = n_qubits; a.k = k; a.conj_t = conj_transpose;
  a.n_tiles = bsz << (n_qubits - TQB_TC_T);
  unsigned long long tmask = 0;
  for (int q = 0; q < k; ++q) {
    if (bitpos_host[q] < 0 || bitpos_host[q] >= n_qubits || ((tmask >> bitpos_host[q]) & 1ull))
      return fail_msg("tqb_apply_dense_tc: bad bit positions");
    tmask |= 1ull << bitpos_host[q];
  }
  unsigned long long tile_mask = tmask;
  for (int p = 0, need = TQB_TC_T - k; need > 0; ++p)
    if (!((tile_mask >> p) & 1ull)) { tile_mask |= 1ull << p; --need; }
  int ti = 0, nshift = 0;
  for (int p = 0; p < n_qubits; ++p) {
    if (!((tile_mask >> p) & 1ull)) continue;
    a.tbit[ti] = p;
    if ((tmask >> p) & 1ull) {
      for (int q = 0; q < k; ++q) if (bitpos_host[q] == p) a.role[ti] = k - 1 - q;
    } else {
      a.role[ti] = -(nshift++) - 1;
    }
    ++ti;
  }
  a.outer_mask = ((n_qubits >= 64 ? ~0ull : (1ull << n_qubits) - 1ull)) & ~tile_mask;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  switch (k) {
    case 4: e = launch_dense_tc<4>(a, g_num_sms, st); break;
    case 5: e = launch_dense_tc<5>(a, g_num_sms, st); break;
    case 6: e = launch_dense_tc<6>(a, g_num_sms, st); break;
    case 7: e = launch_dense_tc<7>(a, g_num_sms, st); break;
  }
  if (e != cudaSuccess) return fail("tqb_dense_tc_kernel", e);
  return 0;
}

int tqb_block_build(const int32_t* blocks_dev, int n_blocks, const int32_t* ops_dev, const float* consts, const float* params_sh,
                    unsigned rank, float* a_planes, void* stream) {
  if (n_blocks <= 0) return 0;
  tqb_block_build_kernel<<<n_blocks, 64, 0, (cudaStream_t)stream>>>(blocks_dev, ops_dev, consts, params_sh, rank, a_planes);
  TQB_CHECK_LAUNCH("tqb_block_build");
  return 0;
}

int tqb_apply_blocks(const float* src, float* dst, float* scratch, int n_qubits, long long bsz, const float* a_planes,
                     const int32_t* blocks_host, int n_blocks, int renorm, void* stream) {
  if (bsz <= 0 || n_blocks <= 0) return 0;
  if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)scratch | (uintptr_t)a_planes) & 15) != 0)
    return fail_msg("tqb_apply_blocks: pointers must be 16-byte aligned");
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  // A rotating block moves the state to the other one of {dst, scratch}, the rest work where the state is; the chain is
  // laid out backwards from "the last block leaves it in dst".  The first block reads the caller's src.
  int flips = 0;
  for (int b = 1; b < n_blocks; ++b) flips += (blocks_host[(size_t)b * TQB_KB_WORDS + TQB_KB_FLAGS] & TQB_KBF_ROTATE) ? 1 : 0;
  const int rot0 = (blocks_host[TQB_KB_FLAGS] & TQB_KBF_ROTATE) ? 1 : 0;
  if ((flips + rot0) > 0 && scratch == nullptr && (flips > 0 || src == dst))
    return fail_msg("tqb_apply_blocks: rotating blocks need a scratch state");
  float* bufs[2] = {dst, scratch};
  int at = flips & 1;   // where the state is after block 0
  const float* cur = src;
  for (int b = 0; b < n_blocks; ++b) {
    const int32_t* bh = blocks_host + (size_t)b * TQB_KB_WORDS;
    int kpos[6];
    for (int k = 0; k < 6; ++k) kpos[k] = bh[TQB_KB_KPOS + k];
    const int rotate = (bh[TQB_KB_FLAGS] & TQB_KBF_ROTATE) ? 1 : 0;
    if (b > 0 && rotate) at ^= 1;
    float* out = bufs[at];
    if (rotate && out == cur)
      return fail_msg("tqb_apply_blocks: with an odd number of rotating blocks dst must differ from src (pass src as scratch)");
    const char* err = tqb_umma::launch_block6(cur, out, n_qubits, bsz, a_planes + (size_t)b * 2 * 128 * 64, kpos,
                                              bh[TQB_KB_FLAGS] & TQB_KBF_TRANSPOSE, rotate, renorm, g_num_sms, (cudaStream_t)stream);
    if (err) return fail_msg(err);
    cur = out;
  }
  return 0;
}

int tqb_quadphase(float* state, int n_local, long long bsz, const uint32_t* coef_dev, int n_bits, unsigned rank, void* stream) {
  if (((uintptr_t)state & 15) != 0) return fail_msg("tqb_quadphase: state must be 16-byte aligned");
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const char* err = tqb_qp::launch_quadphase(state, n_local, bsz, coef_dev, n_bits, rank, g_num_sms, (cudaStream_t)stream);
  if (err) return fail_msg(err);
  return 0;
}

int tqb_dense_grad(const float* psi_in, const float* lam_out, int n_qubits, long long bsz, int k,
                   const int* bitpos_dev, float* G, void* stream) {
  if (bsz <= 0) return 0;
  if (k > 10) return fail_msg("tqb_dense_grad: k too large");
  const long long blocks = bsz << (2 * k);
  tqb_dense_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(psi_in), reinterpret_cast<const float2*>(lam_out), n_qubits, k, bitpos_dev,
      reinterpret_cast<float2*>(G));
  TQB_CHECK_LAUNCH("tqb_dense_grad_kernel");
  return 0;
}

static int expval_chunk_log2(int n) { return n < 12 ? n : 12; }

static int tqb_expval_pauli_slice(const float2* st2, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                                  const unsigned long long* zmask_dev, const int* ny_dev, int n_terms, float* out, int cl,
                                  cudaStream_t st);

int tqb_expval_pauli(const float* state, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                     const unsigned long long* zmask_dev, const int* ny_dev, int n_terms, float* out, void* stream) {
  if (bsz <= 0 || n_terms <= 0) return 0;
  const int cl = expval_chunk_log2(n_qubits);
  const float2* st2 = reinterpret_cast<const float2*>(state);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_qubits >= 1 && n_qubits <= 13 && bsz >= 64) {  // one warp per sample
    const unsigned blocks = (unsigned)((bsz * 32 + 255) / 256);
    if (n_terms == 1)
      tqb_expval_pauli_warp_kernel<1><<<blocks, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, bsz);
    else if (n_terms <= 4)
      tqb_expval_pauli_warp_kernel<4><<<blocks, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, bsz);
    else
      tqb_expval_pauli_warp_kernel<TQB_TERM_BATCH><<<blocks, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, bsz);
    TQB_CHECK_LAUNCH("tqb_expval_pauli_warp_kernel");
    return 0;
  }
  // block-per-chunk kernel: the sample index is gridDim.y (<= 65535); larger batches go in slices
  for (long long b0 = 0; b0 < bsz; b0 += 65535) {
    const long long nb = bsz - b0 < 65535 ? bsz - b0 : 65535;
    const int rc = tqb_expval_pauli_slice(st2 + (b0 << n_qubits), n_qubits, nb, xmask_dev, zmask_dev, ny_dev, n_terms,
                                          out + b0 * n_terms, cl, st);
    if (rc) return rc;
  }
  return 0;
}

static int tqb_expval_pauli_slice(const float2* st2, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                                  const unsigned long long* zmask_dev, const int* ny_dev, int n_terms, float* out, int cl,
                                  cudaStream_t st) {
  dim3 grid(1u << (n_qubits - cl), (unsigned)bsz);
  if (n_terms == 1)
    tqb_expval_pauli_kernel<1><<<grid, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, cl, grid.x > 1);
  else if (n_terms <= 4)
    tqb_expval_pauli_kernel<4><<<grid, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, cl, grid.x > 1);
  else
    tqb_expval_pauli_kernel<TQB_TERM_BATCH><<<grid, 256, 0, st>>>(st2, n_qubits, xmask_dev, zmask_dev, ny_dev, n_terms, out, cl, grid.x > 1);
  TQB_CHECK_LAUNCH("tqb_expval_pauli_kernel");
  return 0;
}

int tqb_expval_pauli_bwd(const float* state, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                         const unsigned long long* zmask_dev, const int* ny_dev, int n_terms, const float* grad_out,
                         float* grad_state, void* stream) {
  const long long total = bsz << n_qubits;
  if (total <= 0) return 0;
  if (n_qubits >= 1 && n_qubits <= 13 && bsz >= 64 && n_terms >= 1) {  // one warp per sample
    tqb_expval_pauli_bwd_warp_kernel<<<(unsigned)((bsz + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(state), n_qubits, bsz, xmask_dev, zmask_dev, ny_dev, n_terms, grad_out,
        reinterpret_cast<float2*>(grad_state));
    TQB_CHECK_LAUNCH("tqb_expval_pauli_bwd_warp_kernel");
    return 0;
  }
  tqb_expval_pauli_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(state), n_qubits, total, xmask_dev, zmask_dev, ny_dev, n_terms, grad_out,
      reinterpret_cast<float2*>(grad_state));
  TQB_CHECK_LAUNCH("tqb_expval_pauli_bwd_kernel");
  return 0;
}

int tqb_expval_z(const float* state, int n_qubits, long long bsz, float* out, void* stream) {
  if (bsz <= 0) return 0;
  const int cl = expval_chunk_log2(n_qubits);
  // the sample index is gridDim.y (<= 65535): larger batches go in slices
  for (long long b0 = 0; b0 < bsz; b0 += 65535) {
    const long long nb = bsz - b0 < 65535 ? bsz - b0 : 65535;
    dim3 grid(1u << (n_qubits - cl), (unsigned)nb);
    tqb_expval_z_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(state) + (b0 << n_qubits), n_qubits,
                                                               out + b0 * (n_qubits + 1), cl, grid.x > 1);
    TQB_CHECK_LAUNCH("tqb_expval_z_kernel");
  }
  return 0;
}

int tqb_expval_z_bwd(const float* state, int n_qubits, long long bsz, const float* grad_out, float* grad_state,
                     void* stream) {
  const long long total = bsz << n_qubits;
  if (total <= 0) return 0;
  tqb_expval_z_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(state), n_qubits, total, grad_out, reinterpret_cast<float2*>(grad_state));
  TQB_CHECK_LAUNCH("tqb_expval_z_bwd_kernel");
  return 0;
}

int tqb_probs(const float* state, int n_qubits, long long bsz, float* probs, void* stream) {
  const long long total = bsz << n_qubits;
  if (total <= 0) return 0;
  tqb_probs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(state), total, probs);
  TQB_CHECK_LAUNCH("tqb_probs_kernel");
  return 0;
}

int tqb_sample_counts(const float* probs, int n_qubits, long long bsz, const float* uniforms, int n_shots,
                      float* cdf_scratch, int* counts, void* stream) {
  if (bsz <= 0) return 0;
  const long long N = 1ll << n_qubits;
  cudaStream_t st = (cudaStream_t)stream;
  tqb_cdf_kernel<<<(unsigned)((bsz + 31) / 32), 32, 0, st>>>(probs, bsz, N, cdf_scratch);
  TQB_CHECK_LAUNCH("tqb_cdf_kernel");
  const long long total = bsz * n_shots;
  if (total > 0) {
    tqb_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cdf_scratch, uniforms, bsz, N, n_shots, counts);
    TQB_CHECK_LAUNCH("tqb_sample_kernel");
  }
  return 0;
}

int tqb_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream) {
  if (n <= 0) return 0;
  if (n % 4 || (((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15)) return fail_msg("tqb_split_tf32: n % 4 == 0 and 16-byte aligned pointers");
  const long long n4 = n / 4;
  const unsigned grid = (unsigned)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  tqb_split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                                              reinterpret_cast<float4*>(lo), n4);
  TQB_CHECK_LAUNCH("tqb_split_tf32_kernel");
  return 0;
}

long long tqb_gemm_workspace_floats(long long M, long long N, long long K) { return tqb_gemm::gemm_workspace_floats(M, N, K); }

int tqb_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* bt_hi, const float* bt_lo, float* c, float* workspace,
                    long long M, long long N, long long K, void* stream) {
  const char* err = tqb_gemm::launch_gemm(a_hi, a_lo, bt_hi, bt_lo, c, workspace, M, N, K, (cudaStream_t)stream);
  if (err) return fail_msg(err);
  return 0;
}

int tqb_gemm_prep(const float* x, long long rows, long long cols, int mode, float* hi, float* lo, void* stream) {
  if (rows <= 0 || cols <= 0 || rows % 32 || cols % 32) return fail_msg("tqb_gemm_prep: rows and cols must be positive multiples of 32");
  if ((((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 7) != 0) return fail_msg("tqb_gemm_prep: pointers must be 8-byte aligned");
  if (cols / 32 > 0x7fffffffll || rows / 32 > 65535) return fail_msg("tqb_gemm_prep: matrix too large");
  dim3 grid((unsigned)(cols / 32), (unsigned)(rows / 32));
  if (mode == 1)
    tqb_gemm_prep_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x), rows, cols,
                                                                   reinterpret_cast<float2*>(hi), reinterpret_cast<float2*>(lo));
  else if (mode == 2)
    tqb_gemm_prep_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x), rows, cols,
                                                                   reinterpret_cast<float2*>(hi), reinterpret_cast<float2*>(lo));
  else
    return fail_msg("tqb_gemm_prep: mode must be 1 (A of x^H) or 2 (Bt of x)");
  TQB_CHECK_LAUNCH("tqb_gemm_prep_kernel");
  return 0;
}

int tqb_exchange_p2p(const float* src, float* const* peer_dst_host, int n_local, long long bsz, int g,
                     const int* victim_bits_host, int rank, void* stream) {
  if (g <= 0 || g > 4 || bsz <= 0) return fail_msg("tqb_exchange_p2p: 1 <= g <= 4 rank-index qubits");
  if (n_local < g + 1 || n_local > 40) return fail_msg("tqb_exchange_p2p: bad n_local");
  ExchangeArgs a;
  a.g = g; a.rank = rank; a.nl = n_local;
  for (int k = 0; k < 4; ++k) a.vpos[k] = k < g ? victim_bits_host[k] : 0;
  for (int k = 0; k < g; ++k) {
    if (a.vpos[k] < 0 || a.vpos[k] >= n_local || (k && a.vpos[k] <= a.vpos[k - 1])) return fail_msg("tqb_exchange_p2p: victim bits must ascend within the slab");
  }
  for (int r = 0; r < 16; ++r) a.dst[r] = r < (1 << g) ? peer_dst_host[r] : nullptr;
  for (int r = 0; r < (1 << g); ++r)
    if (!a.dst[r] || ((uintptr_t)a.dst[r] & 15)) return fail_msg("tqb_exchange_p2p: destination slabs must be 16-byte aligned");
  if ((uintptr_t)src & 15) return fail_msg("tqb_exchange_p2p: source slab must be 16-byte aligned");
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const long long total = bsz << n_local;
  // Large contiguous runs (lowest victim bit high): the copy engines move them at 0.87 of the NVLink rate (777-781 GB/s
  // per direction measured with peer copies on this pool, tools/p2p_probe.py) against 0.70-0.77 for the store kernel,
  // and leave the SMs alone: one cudaMemcpyAsync per run of 2^vpos[0] amplitudes, peers first.
  static int use_ce = -1;
  if (use_ce < 0) { const char* e = getenv("TQB200_EXCHANGE_CE"); use_ce = e ? atoi(e) : 1; }
  const long long run_amps = 1ll << a.vpos[0];
  const long long n_runs = total >> a.vpos[0];
  // (use_ce: 0 never, 1 between two ranks only, 2 for any number of ranks)
  if ((use_ce >= 2 || (use_ce == 1 && g == 1)) && run_amps * 8 >= (8ll << 20) && n_runs <= 1024) {
    unsigned long long vmask = 0;
    for (int k = 0; k < g; ++k) vmask |= 1ull << a.vpos[k];
    const unsigned long long slab_mask = (1ull << n_local) - 1ull;
    // the local part runs on a side stream next to the peer copies (forked from / joined to the caller's stream)
    static cudaStream_t side[16] = {};
    static cudaEvent_t ev_fork[16] = {}, ev_join[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail_msg("tqb_exchange_p2p: device index");
    if (!side[dev]) {
      if (cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming) != cudaSuccess)
        return fail_msg("tqb_exchange_p2p: side stream");
    }
    cudaEventRecord(ev_fork[dev], (cudaStream_t)stream);
    cudaStreamWaitEvent(side[dev], ev_fork[dev], 0);
    // destinations in the order rank + 1, rank + 2, ... (mod 2^g): at any time every rank receives from ONE sender (with
    // the same order on all ranks they would all store into rank 0 first: 7-to-1 incast, 200 GB/s per sender on 8 GPUs);
    // peer destinations on the caller's stream, the local one (step 0, last) on the side stream
    const int W = 1 << g;
    for (int step = 1; step <= W; ++step)
      for (long long r = 0; r < n_runs; ++r) {
        const int want = (rank + step) % W;
        const int pass = want == rank ? 1 : 0;
        const unsigned long long src_idx = (unsigned long long)r << a.vpos[0];   // first amplitude of the run (flat, incl. batch)
        unsigned long long i = src_idx & slab_mask;
        int to = 0;
        for (int k = 0; k < g; ++k) {
          to |= (int)((i >> a.vpos[k]) & 1ull) << k;
          i = (i & ~(1ull << a.vpos[k])) | ((unsigned long long)((rank >> k) & 1) << a.vpos[k]);
        }
        if (to != want) continue;
        const unsigned long long dst_idx = (src_idx & ~slab_mask) | i;
        if (cudaMemcpyAsync(a.dst[to] + 2 * dst_idx, src + 2 * src_idx, (size_t)run_amps * 8, cudaMemcpyDeviceToDevice,
                            pass == 1 ? side[dev] : (cudaStream_t)stream) != cudaSuccess)
          return fail_msg("tqb_exchange_p2p: peer copy failed");
      }
    cudaEventRecord(ev_join[dev], side[dev]);
    cudaStreamWaitEvent((cudaStream_t)stream, ev_join[dev], 0);
    return 0;
  }
  const unsigned grid = (unsigned)g_num_sms * 8u;
  if (a.vpos[0] >= 1) {
    tqb_exchange_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(src, a, total / 2);
  } else {
    tqb_exchange_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(src, a, total);
  }
  TQB_CHECK_LAUNCH("tqb_exchange_kernel");
  return 0;
}

int tqb_host_uniforms(uint32_t* mt_state625_host, long long n, float* out_host) {
  if (n < 0 || !mt_state625_host || (n > 0 && !out_host)) return fail_msg("tqb_host_uniforms: bad arguments");
  if (mt_state625_host[624] > 624u) return fail_msg("tqb_host_uniforms: corrupt generator position");
  tqb_mt19937_uniforms(mt_state625_host, n, out_host);
  return 0;
}

// debug hook (not part of the ABI header): device buffer of 64 x 10 clock64 stamps written by CTA 0 of every following
// tqb_block6_kernel launch; null switches the trace off
void tqb_debug_block_trace(unsigned long long* dev_ptr) { tqb_umma::g_block_trace = dev_ptr; }

}  // extern "C"
