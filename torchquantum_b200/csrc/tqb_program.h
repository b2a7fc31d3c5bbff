// Gate-program wire format shared by the host planner (torchquantum_b200/plan.py), the CUDA
// kernels (tqb_cuda.cu) and the host emulation build used by the CPU tests (tests/emu).
//
// A *pass* is one trip of the state through HBM: every CTA stages one tile of 2^T amplitudes in
// shared memory (T tile bits = the L lowest index bits plus any T-L other bit positions), applies
// the pass's gate list to it and writes it back.  The gate list is cut into *segments* (what is
// staged in shared memory at once: op words + gate matrices [+ gradient accumulators]) and each
// segment into *groups*: a group fixes R "register bits"; every thread holds the 2^R amplitudes
// that differ in those bits and applies all ops of the group to them without touching shared
// memory in between.
//
// All words are int32.  Offsets are in words from the start of the program.
#pragma once
#include <stdint.h>

#define TQB_MAGIC 0x54514231  // "TQB1"
#define TQB_MAX_T 14          // largest tile (2^14 amplitudes = 128 KiB)
#define TQB_MAX_R 4           // register bits per thread (16 amplitudes)
#define TQB_OP_WORDS 10
#define TQB_GROUP_WORDS 8
#define TQB_SEG_HDR_WORDS 4
#define TQB_SEG_MAX_OPS 64
#define TQB_SEG_MAX_GROUPS 64
#define TQB_SEG_MAX_FLOATS 768  // matrix floats staged per segment (per sample)
#define TQB_SEG_MAX_WORDS (TQB_SEG_HDR_WORDS + TQB_SEG_MAX_GROUPS * TQB_GROUP_WORDS + TQB_SEG_MAX_OPS * TQB_OP_WORDS)

// program header words
#define TQB_H_MAGIC 0
#define TQB_H_T 1
#define TQB_H_R 2
#define TQB_H_NSEG 3
#define TQB_H_L 4          // number of contiguous low bits in the tile
#define TQB_H_WORDS 5      // total program length
#define TQB_H_OUTER 6      // number of non-tile bits (n - T)
#define TQB_H_TBIT 8       // 16 words: global bit position of tile-local bit i
#define TQB_H_SEGOFF 24    // n_seg words: offset of each segment

// segment header: [n_groups, n_ops, mat_base, mat_len]; then n_groups group records; then ops.
// group record: [regpos0..3, n_ops, first_op, 0, 0]  (regpos = tile-local bit of register bit j)
// op record (10 words):
#define TQB_OP_TYPE 0   // low 8 bits type, bits 8.. flags
#define TQB_OP_TGT 1    // U1Q: register-bit index j.  U2Q: unused (targets are register bits 1,0)
#define TQB_OP_PM_T 2   // DIAG: parity mask over tile-local bits
#define TQB_OP_CM_T 3   // control mask over tile-local bits
#define TQB_OP_CV_T 4   // required value of the controlled bits
#define TQB_OP_SLOT 5   // float offset of this op's matrix inside the segment's staged matrices
#define TQB_OP_PM_O 6   // DIAG: parity mask over the compressed outer (non-tile) index
#define TQB_OP_CM_O 7   // control mask over the compressed outer index
#define TQB_OP_CV_O 8
#define TQB_OP_AUX 9

#define TQB_T_U1Q 0   // (controlled) 2x2 on one register bit; 8 floats row-major (re,im)
#define TQB_T_DIAG 1  // (controlled) amp *= d[parity(idx & pm)]; 4 floats d0,d1
#define TQB_T_U2Q 2   // (controlled) 4x4 on register bits (1,0) = (msb,lsb); 32 floats

#define TQB_F_GRAD 0x100     // accumulate dL/dU for this op in the backward sweep
#define TQB_F_NOUNAPPLY 0x200  // backward: psi already holds the op's INPUT (non-unitary single-op node)

// matrix-builder table: one record of 8 words per op (tqb_build_mats / tqb_build_mats_bwd)
#define TQB_B_WORDS 8
#define TQB_B_KIND 0      // builder kind (below)
#define TQB_B_FLAGS 1     // bit0 inverse, bit1 per-sample params
#define TQB_B_POFF 2      // first parameter column (in params_sh or params_ps)
#define TQB_B_MOFF 3      // float offset in the matrix row
#define TQB_B_COFF 4      // K_CONST: float offset into the constant pool
#define TQB_B_NFLOAT 5    // floats written (8, 4 or 32)
#define TQB_B_NPARAM 6

#define TQB_BF_INVERSE 1
#define TQB_BF_PER_SAMPLE 2

enum {
  TQB_K_CONST = 0,   // copy NFLOAT floats from the constant pool (inverse already folded in)
  TQB_K_EXTERN = 1,  // matrix written by the host front end (explicit unitary); builder skips it
  TQB_K_RX = 2, TQB_K_RY = 3, TQB_K_ROT = 4, TQB_K_U2 = 5, TQB_K_U2RAW = 6, TQB_K_U3 = 7,
  TQB_K_CU = 8, TQB_K_R = 9,                     // -> 2x2 (8 floats)
  TQB_K_RZ = 10, TQB_K_PHASE = 11, TQB_K_GPHASE = 12,  // -> diagonal pair d0,d1 (4 floats)
  TQB_K_RXX = 13, TQB_K_RYY = 14, TQB_K_RZX = 15, TQB_K_SEXC = 16, TQB_K_XXMYY = 17,
  TQB_K_XXPYY = 18,                               // -> 4x4 (32 floats)
  TQB_K_COUNT = 19
};
