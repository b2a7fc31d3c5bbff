// On-device block-unitary builder (single source: nvcc for sm_100a, g++ for the CPU tests' emulation build).
//
// Replaces the host-side block products of the reference's static mode, torchquantum/graph/graphs.py:713-861
// (`get_unitary` of each block of <= wires_per_block qubits, multiplied with torch.matmul on full 2^w x 2^w
// matrices), for the forward-only block programs of csrc/tqb_program.h: the gates of one block are applied, in
// program order, to the 64 basis columns of a 6-qubit identity (one thread per column, double precision), which
// leaves column j of the block unitary U; U is then written as the two TF32 planes (hi, lo) of
// A[2i + w][j] = (w ? Im : Re) U[i][j] that tqb_block6_kernel feeds to the tensor cores.
// Gate matrices come from the same build_gate<T> as the pass kernels' matrix builder, instantiated for double.
#pragma once
#include "tqb_core.cuh"

namespace tqb {

typedef Cx<double> cxd;
TQB_FN cxd cxd_mul(cxd a, cxd b) { return cxd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
TQB_FN cxd cxd_add(cxd a, cxd b) { return cxd{a.re + b.re, a.im + b.im}; }

// round-to-nearest (ties away) to TF32, as cvt.rna.tf32.f32 does
TQB_FN float tf32_round(float x) {
  u32 b;
#if defined(__CUDA_ARCH__)
  b = __float_as_uint(x);
#else
  memcpy(&b, &x, 4);
#endif
  b = (b + 0x1000u) & ~0x1fffu;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float r;
  memcpy(&r, &b, 4);
  return r;
#endif
}

// the op's applied matrix (inverse folded in) in full complex form: 4 / 2 / 16 entries for U1 / DIAG / U2
TQB_FN void block_op_matrix(const int32_t* op, const float* consts, const float* params_sh, cxd* m) {
  const u32 w0 = (u32)op[TQB_KO_W0];
  const int form = (int)(w0 & 15u), np = (int)((w0 >> 4) & 15u);
  const bool inv = (w0 >> 8) & 1u;
  const int kind = op[TQB_KO_KIND];
  const int nent = form == TQB_KF_U1 ? 4 : (form == TQB_KF_DIAG ? 2 : (form == TQB_KF_U2 ? 16 : 0));
  if (kind == TQB_K_CONST) {
    const float* c = consts + op[TQB_KO_OFF];
    for (int i = 0; i < nent; ++i) m[i] = cxd{(double)c[2 * i], (double)c[2 * i + 1]};
    return;
  }
  double th[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < np; ++i) th[i] = (double)params_sh[op[TQB_KO_OFF] + i];
  cxd u[16];
  const int got = build_gate<double>(kind, th, u);
  const int d = kind_dim(got);
  if (d == 0) {
    for (int i = 0; i < 2; ++i) m[i] = cxd{u[i].re, inv ? -u[i].im : u[i].im};
    return;
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      const cxd v = inv ? u[j * d + i] : u[i * d + j];
      m[i * d + j] = cxd{v.re, inv ? -v.im : v.im};
    }
}

// applies the block's ops to one column (64 amplitudes over the block-local bit patterns i)
TQB_FN void block_build_column(const int32_t* ops, int nops, const float* consts, const float* params_sh, u32 rank, cxd* col) {
  for (int k = 0; k < nops; ++k) {
    const int32_t* op = ops + k * TQB_KO_WORDS;
    if ((rank & (u32)op[TQB_KO_RCM]) != (u32)op[TQB_KO_RCV]) continue;
    const int form = op[TQB_KO_W0] & 15;
    const u32 t0 = (u32)op[TQB_KO_BITS] & 255u, t1 = ((u32)op[TQB_KO_BITS] >> 8) & 255u;
    const u32 cm = (u32)op[TQB_KO_MASKS] & 255u, cv = ((u32)op[TQB_KO_MASKS] >> 8) & 255u, pm = ((u32)op[TQB_KO_MASKS] >> 16) & 255u;
    cxd m[16];
    block_op_matrix(op, consts, params_sh, m);
    if (form == TQB_KF_DIAG) {
      const u32 rp = (u32)__popc(rank & (u32)op[TQB_KO_RPM]) & 1u;
      for (u32 i = 0; i < 64u; ++i)
        if ((i & cm) == cv) col[i] = cxd_mul(m[((u32)__popc(i & pm) & 1u) ^ rp], col[i]);
    } else if (form == TQB_KF_X) {
      for (u32 i = 0; i < 64u; ++i)
        if (!((i >> t0) & 1u) && (i & cm) == cv) {
          const cxd a = col[i];
          col[i] = col[i | (1u << t0)];
          col[i | (1u << t0)] = a;
        }
    } else if (form == TQB_KF_U1) {
      for (u32 i = 0; i < 64u; ++i)
        if (!((i >> t0) & 1u) && (i & cm) == cv) {
          const cxd a = col[i], b = col[i | (1u << t0)];
          col[i] = cxd_add(cxd_mul(m[0], a), cxd_mul(m[1], b));
          col[i | (1u << t0)] = cxd_add(cxd_mul(m[2], a), cxd_mul(m[3], b));
        }
    } else {  // TQB_KF_U2: quad index = 2 * bit(t1) + bit(t0)
      for (u32 i = 0; i < 64u; ++i)
        if (!((i >> t0) & 1u) && !((i >> t1) & 1u) && (i & cm) == cv) {
          const u32 idx[4] = {i, i | (1u << t0), i | (1u << t1), i | (1u << t0) | (1u << t1)};
          cxd v[4], o[4];
          for (int q = 0; q < 4; ++q) v[q] = col[idx[q]];
          for (int r = 0; r < 4; ++r) {
            o[r] = cxd{0.0, 0.0};
            for (int q = 0; q < 4; ++q) o[r] = cxd_add(o[r], cxd_mul(m[4 * r + q], v[q]));
          }
          for (int q = 0; q < 4; ++q) col[idx[q]] = o[q];
        }
    }
  }
}

// column j of U -> column j of the two A planes ([2][128][64] floats per block)
TQB_FN void block_store_column(const cxd* col, int j, float* planes) {
  for (int i = 0; i < 64; ++i)
    for (int w = 0; w < 2; ++w) {
      const double v = w ? col[i].im : col[i].re;
      const float hi = tf32_round((float)v);
      const float lo = tf32_round((float)(v - (double)hi));
      planes[(size_t)(2 * i + w) * 64 + j] = hi;
      planes[(size_t)(128 + 2 * i + w) * 64 + j] = lo;
    }
}

// one (block, column)
TQB_FN void block_build_one(const int32_t* blocks, const int32_t* ops, const float* consts, const float* params_sh, u32 rank,
                            int block, int j, float* planes_all) {
  const int32_t* bh = blocks + (size_t)block * TQB_KB_WORDS;
  cxd col[64];
  for (int i = 0; i < 64; ++i) col[i] = cxd{i == j ? 1.0 : 0.0, 0.0};
  block_build_column(ops + (size_t)bh[TQB_KB_OPFIRST] * TQB_KO_WORDS, bh[TQB_KB_NOPS], consts, params_sh, rank, col);
  block_store_column(col, j, planes_all + (size_t)block * 2 * 128 * 64);
}

}  // namespace tqb
