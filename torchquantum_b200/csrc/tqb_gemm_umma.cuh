// Real GEMM  C[M][N] = A[M][K] * Bt[N][K]^T  in fp32 accuracy on the Blackwell tensor cores (tcgen05.mma kind::tf32 with the
// 3xTF32 split: C = Alo Bhi + Ahi Blo + Ahi Bhi, fp32 accumulation in TMEM).
//
// Used by the engine's batch collapse (engine.py: a run of gates shared by the whole batch is ONE 2^n x 2^n unitary W, built
// on the identity batch and applied to all bsz states at once).  The complex product  out = psi W^T  is this real GEMM on the
// interleaved (re, im) views: A = psi as [bsz][2 * 2^n], Bt[2i + w][2c + v] = the real 2x2 embedding of W^T[c][i], and the
// same kernel serves the two GEMMs of the backward.  It replaces what the reference's static mode does with one torch.bmm
// per fused block (torchquantum/graph/graphs.py:836-861, QuantumGraph.apply_unitary).
//
// Persistent CTAs (one per SM) walk the 128 x 256 tiles of C: warp 0 streams 128-byte K slices of the operand planes with TMA (2-D tensor maps,
// SWIZZLE_128B, 4 stages of 16 + 32 KiB), warp 1 issues 4 MMAs (M = 128, N = 256, K = 8) per slice into a 128 x 256 fp32
// accumulator in TMEM (two of them: the epilogue of a tile overlaps the MMAs of the next), warps 4-7 read it back
// (tcgen05.ld 32x32b) and store 128 contiguous bytes per thread and load.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tqb_block_umma.cuh"   // mbarrier / TMA / tcgen05 helpers, tensor-map encoder

namespace tqb_gemm {

using tqb_umma::mbar_arrive;
using tqb_umma::mbar_expect_tx;
using tqb_umma::mbar_init;
using tqb_umma::mbar_wait;
using tqb_umma::smem_u32;
using tqb_umma::umma_commit;

constexpr int BM = 128, BN = 256, BK = 32;       // BK floats = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr uint32_t kABytes = BM * BK * 4, kBBytes = BN * BK * 4, kStage = kABytes + kBBytes;
constexpr uint32_t kSmem = kStages * kStage + 1024 + 256;
constexpr int kThreads = 256;
// instruction descriptor: D fp32, A / B tf32, both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// K-major operand tile, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO = 1 (unused)
__device__ __forceinline__ uint64_t k_desc(uint32_t addr) {
  const uint32_t lo = ((addr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

struct GemmParams {
  float* c;              // [splits][M][N] partial products (splits == 1: the result itself)
  long long split_stride;  // M * N
  int ldc;               // = N
  int k_slices;          // K slices of BK floats handled by one tile (one split)
  int a_hi_plane, b_hi_plane;   // which plane (0 / 1) of the operand's tensor map is its TF32 hi part
  int tiles_n, tiles_m, n_tiles;   // tile grid: tile = (split * tiles_m + m) * tiles_n + n
};

constexpr uint32_t kAccs = 2;   // TMEM accumulators (BN columns each): the epilogue of a tile overlaps the next tile's MMAs

// Persistent: one CTA per SM walks the tiles (stride gridDim.x); the TMA / MMA pipelines run on across tile boundaries.
__global__ void __launch_bounds__(kThreads, 1)
tqb_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmParams p) {
  // ONE 3-D tensor map per operand (K, rows, {hi, lo} plane): four maps made the parameter block 536 bytes, and kernels
  // whose parameter block is longer than 384 bytes run slower on B200 (measured on tqb_block6_kernel)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kStages * kStage;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto accfull = [&](uint32_t a) { return bars + 8u * (2 * kStages + a); };
  auto accempty = [&](uint32_t a) { return bars + 8u * (2 * kStages + kAccs + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 2 * kAccs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (uint32_t a = 0; a < kAccs; ++a) { mbar_init(accfull(a), 1); mbar_init(accempty(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(kAccs * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

  const int total = 3 * p.k_slices;   // pass 0: Alo Bhi, pass 1: Ahi Blo, pass 2: Ahi Bhi (the small products first: the
                                      // accumulator is still tiny while the truncating fp32 adds of the tensor core hit them)
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;       // slices issued by this CTA so far (the stage ring runs across tiles)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.tiles_n) * BN, m0 = ((tile / p.tiles_n) % p.tiles_m) * BM;
        const int ks0 = (tile / (p.tiles_n * p.tiles_m)) * p.k_slices;   // split-K: this tile's share of the K slices
        for (int i = 0; i < total; ++i, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          const int pass = i / p.k_slices, ks = i - pass * p.k_slices;
          mbar_wait(empty(s), ph ^ 1u);
          mbar_expect_tx(full(s), kStage);
          const uint32_t sa = base + s * kStage, sb = sa + kABytes;
          tma_load_3d(sa, &tm_a, full(s), (ks0 + ks) * BK, m0, pass == 0 ? 1 - p.a_hi_plane : p.a_hi_plane);
          tma_load_3d(sb, &tm_b, full(s), (ks0 + ks) * BK, n0, pass == 1 ? 1 - p.b_hi_plane : p.b_hi_plane);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;   // slices / tiles done by this CTA
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt % kAccs;
        mbar_wait(accempty(acc), ((lt / kAccs) & 1u) ^ 1u);    // the epilogue has read this accumulator's previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + acc * (uint32_t)BN;
        for (int i = 0; i < total; ++i, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(full(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * kStage, sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)   // K = 8 per MMA: 32 bytes further along the swizzled 128-byte row
            umma_tf32_ss(d, k_desc(sa + 32u * k), k_desc(sb + 32u * k), (i | k) != 0);
          umma_commit(empty(s));
        }
        umma_commit(accfull(acc));
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
      const uint32_t acc = lt % kAccs;
      const int n0 = (tile % p.tiles_n) * BN, m0 = ((tile / p.tiles_n) % p.tiles_m) * BM;
      const int split = tile / (p.tiles_n * p.tiles_m);
      mbar_wait(accfull(acc), (lt / kAccs) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* crow = p.c + (size_t)split * p.split_stride + (size_t)(m0 + 32 * q + lane) * p.ldc + n0;
#pragma unroll 1
      for (int g = 0; g < BN / 32; ++g) {
        uint32_t v[32];
        tqb_umma::tmem_ld32(tmem + acc * (uint32_t)BN + ((uint32_t)(32 * q) << 16) + 32u * g, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4* dst = reinterpret_cast<float4*>(crow + 32 * g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(accempty(acc));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(kAccs * BN)) : "memory");
  }
}

// 3-D tensor map over the two planes (TF32 hi / lo parts, anywhere in memory) of a row-major [rows][K] fp32 matrix:
// dims (K, rows, 2), box = BK floats x box_rows x 1, SWIZZLE_128B.  *hi_plane: the plane index of `hi`.
inline const char* make_map_planes(CUtensorMap* tm, const float* hi, const float* lo, long long rows, long long K, int box_rows,
                                   int* hi_plane) {
  tqb_umma::EncodeTiledFn enc = tqb_umma::encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled not available";
  const float* ptr = hi < lo ? hi : lo;
  *hi_plane = hi < lo ? 0 : 1;
  const unsigned long long gap = (unsigned long long)((hi < lo ? lo : hi) - ptr) * 4ull;
  if (gap == 0 || (gap & 15ull) || gap >= (1ull << 40)) return "gemm: the hi / lo planes of an operand must be distinct and within 2^40 bytes";
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)gap};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled failed (gemm)";
}

// c[i] = sum_z ws[z][i] in round-to-nearest fp32 (the second level of the split-K accumulation)
__global__ void __launch_bounds__(256) tqb_gemm_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ c, long long n4, int splits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = ws[i];
    for (int z = 1; z < splits; ++z) {
      const float4 v = ws[(long long)z * n4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    c[i] = acc;
  }
}

// Split-K: the tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of the chain
// (measured: 4.8e-6 relative after K = 2048, 1.9e-5 after K = 8192 in one accumulator).  Chains are therefore cut at
// kMaxChain floats of K; the partial products go to a workspace and are summed in round-to-nearest by a second kernel
// (which also fills the machine when the tile grid alone is smaller than the GPU).
constexpr long long kMaxChain = 512;

inline long long gemm_splits(long long K) {
  long long s = (K + kMaxChain - 1) / kMaxChain;
  while (s > 1 && (K / BK) % s) --s;   // whole slices per split
  return s < 1 ? 1 : s;
}

// workspace floats needed by launch_gemm (0: no split)
inline long long gemm_workspace_floats(long long M, long long N, long long K) {
  const long long s = gemm_splits(K);
  return s > 1 ? s * M * N : 0;
}

inline const char* launch_gemm(const float* a_hi, const float* a_lo, const float* bt_hi, const float* bt_lo, float* c, float* workspace,
                               long long M, long long N, long long K, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0 || M % BM || N % BN || K % BK) return "gemm: M % 128, N % 256, K % 32 must be 0";
  if ((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)bt_hi | (uintptr_t)bt_lo | (uintptr_t)c | (uintptr_t)workspace) & 15) != 0)
    return "gemm: pointers must be 16-byte aligned";
  static int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != configured_dev) {
    if (cudaFuncSetAttribute(tqb_gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess)
      return "cudaFuncSetAttribute(gemm) failed";
    configured_dev = dev;
  }
  const long long splits = gemm_splits(K);
  if (splits > 1 && !workspace) return "gemm: workspace of splits * M * N floats needed";
  CUtensorMap ta, tb;
  const char* err;
  GemmParams p;
  if ((err = make_map_planes(&ta, a_hi, a_lo, M, K, BM, &p.a_hi_plane))) return err;
  if ((err = make_map_planes(&tb, bt_hi, bt_lo, N, K, BN, &p.b_hi_plane))) return err;
  p.c = splits > 1 ? workspace : c;
  p.split_stride = M * N;
  p.ldc = (int)N;
  p.k_slices = (int)(K / BK / splits);
  p.tiles_n = (int)(N / BN);
  p.tiles_m = (int)(M / BM);
  const long long n_tiles = (N / BN) * (M / BM) * splits;
  if (n_tiles > 0x7fffffffll) return "gemm: too many tiles";
  p.n_tiles = (int)n_tiles;
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  tqb_gemm_tf32x3_kernel<<<grid, kThreads, kSmem, st>>>(ta, tb, p);
  if (cudaGetLastError() != cudaSuccess) return "gemm launch failed";
  if (splits > 1) {
    const long long n4 = M * N / 4;
    const unsigned rgrid = (unsigned)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    tqb_gemm_reduce_kernel<<<rgrid, 256, 0, st>>>(reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(c), n4, (int)splits);
    if (cudaGetLastError() != cudaSuccess) return "gemm reduce launch failed";
  }
  return nullptr;
}

}  // namespace tqb_gemm
