// K4: dense k-qubit unitary (4 <= k <= 7) on the tensor cores, 3xTF32 error-compensated.
//
//   out[b, i, rest] = sum_j U[i, j] in[b, j, rest]        i, j: patterns of the k target bits (matrix MSB first)
//
// Every CTA loops over tiles of 2^12 amplitudes = the k target bits + the 12-k lowest other bits, stages a tile with
// cp.async one tile ahead as B[n][j] (n: pattern of the non-target tile bits), and computes the complex product
//   Cr = Ur Br - Ui Bi,   Ci = Ur Bi + Ui Br
// with mma.sync m16n8k8 TF32 MMAs.  Each fp32 operand x is split as hi = tf32(x), lo = tf32(x - hi) and a product uses
// hi*hi + hi*lo + lo*hi (the lo*lo term is below fp32 rounding), which keeps the result within ~1e-6 of an fp32
// contraction: SURVEY §8c leaves no room for a plain TF32 path under the 1e-5 state tolerance.
// Warp w owns the 16-row strip (w % strips) of C and one column group: 8 warps = strips x column groups.
// (mma.sync, not tcgen05: this is the first tensor-core form of K4; it serves tqf.qft / qubitunitary* on 4-7 wires and
// measures what a dense block costs next to the gate-by-gate interpreter.  DESIGN.md §3.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tqb {

#define TQB_TC_T 12          // tile bits
#define TQB_TC_THREADS 256

struct DenseTcArgs {
  const float2* in;
  float2* out;
  const float2* mat;      // [M][M] complex64, row-major
  int n, k, conj_t;
  long long n_tiles;      // bsz << (n - T)
  int tbit[TQB_TC_T];     // global bit position of tile-local bit i (ascending)
  int role[TQB_TC_T];     // tile-local bit i: >= 0: matrix index bit (value = shift in the row/col index); < 0: -(shift in n)-1
  unsigned long long outer_mask;  // bits of the amplitude index that are NOT tile bits
};

__device__ __forceinline__ uint32_t tc_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tc_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tc_tf32(x);
  lo = tc_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void tc_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += a * b with the 3xTF32 split (a, b given as hi / lo fragments)
__device__ __forceinline__ void tc_mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                        const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  tc_mma(c, al, bh);
  tc_mma(c, ah, bl);
  tc_mma(c, ah, bh);
}

__device__ __forceinline__ void tc_cp_async8(float2* dst, const float2* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}

template <int K>
__global__ void __launch_bounds__(TQB_TC_THREADS, (K <= 6 ? 2 : 1)) tqb_dense_tc_kernel(const DenseTcArgs a) {
  constexpr int M = 1 << K;                 // matrix dimension
  constexpr int N = 1 << (TQB_TC_T - K);    // columns per tile
  constexpr int LD = M + 4;                 // leading dimension of every plane (bank-conflict-free fragments)
  constexpr int STRIPS = M / 16;            // 16-row strips of C
  constexpr int CGROUPS = 8 / STRIPS;       // column groups
  constexpr int NT = N / CGROUPS / 8;       // 8-column MMA tiles per warp
  constexpr int PER = (1 << TQB_TC_T) / TQB_TC_THREADS;   // amplitudes a thread moves per tile
  static_assert(NT >= 1 && STRIPS * CGROUPS == 8, "warp layout");
  // k <= 5: U is split into its TF32 hi / lo parts once per CTA (four planes); k >= 6: two fp32 planes, split at every use
  // (four planes would cost the second resident CTA at k = 6 and do not fit at k = 7)
  constexpr bool PRESPLIT = K <= 5;
  constexpr int UPLANES = PRESPLIT ? 4 : 2;
  extern __shared__ __align__(16) float smem_tc[];
  float* Ur = smem_tc;                      // [M][LD]   (PRESPLIT: hi part as tf32 bits)
  float* Ui = Ur + M * LD;
  float* Url = Ui + M * LD;                 // PRESPLIT only: lo parts
  float* Uil = Url + M * LD;
  // two tile buffers B[n][j] of complex amplitudes (column n, matrix index j contiguous), filled by cp.async one tile
  // ahead; the product C overwrites the buffer it was computed from and is written back from there
  float2* Bbuf = reinterpret_cast<float2*>(smem_tc + UPLANES * M * LD);
  // tile-local index = lo6 | hi6 << 6: two 64-entry tables give its in-tile amplitude offset and (j, n) coordinates
  __shared__ unsigned long long s_off[128];
  __shared__ uint32_t s_jn[128];          // j | n << 16
  __shared__ unsigned long long s_base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 128) {
    const int half = tid >> 6;
    const uint32_t idx = (uint32_t)(tid & 63) << (6 * half);
    unsigned long long off = 0;
    uint32_t j = 0, nn = 0;
    for (int bpos = 0; bpos < TQB_TC_T; ++bpos) {
      const uint32_t bit = (idx >> bpos) & 1u;
      off |= (unsigned long long)bit << a.tbit[bpos];
      if (a.role[bpos] >= 0) j |= bit << a.role[bpos]; else nn |= bit << (-a.role[bpos] - 1);
    }
    s_off[tid] = off;
    s_jn[tid] = j | (nn << 16);
  }
  // U (or U^H) into the planes, once per CTA
  for (int e = tid; e < M * M; e += TQB_TC_THREADS) {
    const int r = e / M, c = e % M;
    const float2 v = a.conj_t ? a.mat[c * M + r] : a.mat[e];
    const float im = a.conj_t ? -v.y : v.y;
    if (PRESPLIT) {
      uint32_t h, l;
      tc_split(v.x, h, l);
      Ur[r * LD + c] = __uint_as_float(h); Url[r * LD + c] = __uint_as_float(l);
      tc_split(im, h, l);
      Ui[r * LD + c] = __uint_as_float(h); Uil[r * LD + c] = __uint_as_float(l);
    } else {
      Ur[r * LD + c] = v.x;
      Ui[r * LD + c] = im;
    }
  }
  const int strip = warp % STRIPS, cgroup = warp / STRIPS;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles_log2 = a.n - TQB_TC_T;
  // amplitude offset of a tile: sample * 2^n + the tile index deposited on the non-tile bits
  auto tile_base = [&](long long tile) {
    const unsigned long long sample = (unsigned long long)tile >> ntiles_log2;
    unsigned long long rest = (unsigned long long)tile & ((1ull << ntiles_log2) - 1ull), bb = sample << a.n;
    unsigned long long m = a.outer_mask;
    while (m) {
      const int p = __ffsll((long long)m) - 1;
      m &= m - 1ull;
      bb |= (rest & 1ull) << p;
      rest >>= 1;
    }
    return bb;
  };
  // this thread's 16 (offset, smem slot) pairs are the same for every tile
  auto issue = [&](unsigned long long base, float2* B) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const uint32_t idx = (uint32_t)tid + TQB_TC_THREADS * i;
      const unsigned long long off = s_off[idx & 63u] | s_off[64u + (idx >> 6)];
      const uint32_t jn = s_jn[idx & 63u] | s_jn[64u + (idx >> 6)];
      tc_cp_async8(B + (jn >> 16) * LD + (jn & 0xffffu), a.in + base + off);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  if (tid == 0 && (long long)blockIdx.x < a.n_tiles) s_base[0] = tile_base(blockIdx.x);
  __syncthreads();
  if ((long long)blockIdx.x < a.n_tiles) issue(s_base[0], Bbuf);
  int buf = 0;
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    float2* B = Bbuf + (size_t)buf * N * LD;
    const long long next = tile + gridDim.x;
    if (tid == 0 && next < a.n_tiles) s_base[buf ^ 1] = tile_base(next);
    __syncthreads();   // s_base[buf ^ 1] visible; the other buffer was drained by the previous iteration's store
    const unsigned long long base = s_base[buf];
    if (next < a.n_tiles) {
      issue(s_base[buf ^ 1], Bbuf + (size_t)(buf ^ 1) * N * LD);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();   // this tile's amplitudes have landed for every thread
    // ---- C strip = U strip x B (complex, 3xTF32)
    float cr[NT][4], ci[NT][4];
#pragma unroll
    for (int q = 0; q < NT; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) { cr[q][e] = 0.f; ci[q][e] = 0.f; }
    const int row0 = strip * 16;
    for (int k0 = 0; k0 < M; k0 += 8) {
      uint32_t urh[4], url[4], uih[4], uil[4], nuih[4], nuil[4];
      {
        const int o = (row0 + g) * LD + k0 + t;
        const int fo[4] = {o, o + 8 * LD, o + 4, o + 8 * LD + 4};   // A fragment: (g, t) (g+8, t) (g, t+4) (g+8, t+4)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (PRESPLIT) {
            urh[e] = __float_as_uint(Ur[fo[e]]); url[e] = __float_as_uint(Url[fo[e]]);
            uih[e] = __float_as_uint(Ui[fo[e]]); uil[e] = __float_as_uint(Uil[fo[e]]);
          } else {
            tc_split(Ur[fo[e]], urh[e], url[e]);
            tc_split(Ui[fo[e]], uih[e], uil[e]);
          }
          nuih[e] = uih[e] ^ 0x80000000u;    // tf32 rounding is symmetric: split(-x) = -split(x)
          nuil[e] = uil[e] ^ 0x80000000u;
        }
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const int col = (cgroup * NT + q) * 8 + g;     // B fragment: (k = t / t + 4, n = g), re and im in one 64-bit read
        const float2 b0 = B[col * LD + k0 + t], b1 = B[col * LD + k0 + t + 4];
        uint32_t brh[2], brl[2], bih[2], bil[2];
        tc_split(b0.x, brh[0], brl[0]); tc_split(b1.x, brh[1], brl[1]);
        tc_split(b0.y, bih[0], bil[0]); tc_split(b1.y, bih[1], bil[1]);
        tc_mma3(cr[q], urh, url, brh, brl);      // Cr += Ur Br
        tc_mma3(cr[q], nuih, nuil, bih, bil);    // Cr -= Ui Bi
        tc_mma3(ci[q], urh, url, bih, bil);      // Ci += Ur Bi
        tc_mma3(ci[q], uih, uil, brh, brl);      // Ci += Ui Br
      }
    }
    __syncthreads();   // every warp is done reading B: reuse the buffer for C
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const int col = (cgroup * NT + q) * 8 + 2 * t;    // C fragment: rows g / g + 8, columns 2t / 2t + 1
      B[col * LD + row0 + g] = make_float2(cr[q][0], ci[q][0]);
      B[(col + 1) * LD + row0 + g] = make_float2(cr[q][1], ci[q][1]);
      B[col * LD + row0 + g + 8] = make_float2(cr[q][2], ci[q][2]);
      B[(col + 1) * LD + row0 + g + 8] = make_float2(cr[q][3], ci[q][3]);
    }
    __syncthreads();
    // ---- store
#pragma unroll 4
    for (int i = 0; i < PER; ++i) {
      const uint32_t idx = (uint32_t)tid + TQB_TC_THREADS * i;
      const unsigned long long off = s_off[idx & 63u] | s_off[64u + (idx >> 6)];
      const uint32_t jn = s_jn[idx & 63u] | s_jn[64u + (idx >> 6)];
      a.out[base + off] = B[(jn >> 16) * LD + (jn & 0xffffu)];
    }
    buf ^= 1;
  }
}

template <int K>
static cudaError_t launch_dense_tc(const DenseTcArgs& a, int num_sms, cudaStream_t st) {
  constexpr int M = 1 << K, N = 1 << (TQB_TC_T - K), LD = M + 4;
  const size_t smem = (size_t)((K <= 5 ? 4 : 2) * M * LD) * sizeof(float) + (size_t)2 * N * LD * sizeof(float2);
  static int configured_dev = -1;   // kernel attributes are per device
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  if (dev_now != configured_dev) {
    cudaError_t e = cudaFuncSetAttribute(tqb_dense_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev_now;
  }
  // k <= 6: two CTAs per SM (<= 103 KB each); k = 7: one (201 KB)
  const long long resident = (long long)num_sms * (K <= 6 ? 2 : 1);
  long long grid = a.n_tiles < resident ? a.n_tiles : resident;
  tqb_dense_tc_kernel<K><<<(unsigned)grid, TQB_TC_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace tqb
