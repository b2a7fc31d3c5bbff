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
// Gate matrices live in two rows: one row shared by the whole batch (constants and gates whose
// parameters are shared, `msh`) and one row per sample (`mps`, per-sample parameters).  Gradient
// accumulators are one row per sample (`gr`).
//
// All words are int32.  Offsets are in words from the start of the program.
#pragma once
#include <stdint.h>

#define TQB_MAGIC 0x54514232  // "TQB2"
#define TQB_MAX_T 13          // largest tile (2^13 amplitudes = 64 KiB)
#define TQB_MAX_R 4           // register bits per thread (16 amplitudes)
#define TQB_OP_WORDS 12
#define TQB_GROUP_WORDS 8
#define TQB_SEG_HDR_WORDS 8
#define TQB_SEG_MAX_OPS 64
#define TQB_SEG_MAX_GROUPS 32
#define TQB_SEG_MAX_SH 768   // shared matrix floats staged per segment
#define TQB_SEG_MAX_PS 256   // per-sample matrix floats staged per segment (per sample)
#define TQB_SEG_MAX_G 48     // gradient floats per segment (per sample): the backward stages [G][threads + 1] partials
#define TQB_SEG_MAX_WORDS (TQB_SEG_HDR_WORDS + TQB_SEG_MAX_GROUPS * TQB_GROUP_WORDS + TQB_SEG_MAX_OPS * TQB_OP_WORDS)

// program header words
#define TQB_H_MAGIC 0
#define TQB_H_T 1
#define TQB_H_R 2
#define TQB_H_NSEG 3
#define TQB_H_L 4          // number of contiguous low bits in the tile
#define TQB_H_WORDS 5      // total program length
#define TQB_H_OUTER 6      // number of non-tile bits (n_total - T)
#define TQB_H_PS_MAX 7     // max per-sample matrix floats of any segment
#define TQB_H_TBIT 8       // 16 words: global bit position of tile-local bit i
#define TQB_H_G_MAX 24     // max gradient floats of any segment
#define TQB_H_SH_MAX 25    // max shared matrix floats of any segment
#define TQB_H_SCALE_SH 26  // float offset (shared matrix row) of the pass's deferred complex scale factor, -1: none
#define TQB_H_SCALE_PS 27  // float offset (per-sample matrix rows) of the per-sample part of that factor, -1: none
#define TQB_H_SEGOFF 32    // n_seg words: offset of each segment

// segment header (8 words); then n_groups group records; then the op records
#define TQB_S_NGROUPS 0
#define TQB_S_NOPS 1
#define TQB_S_SH_BASE 2   // first float of this segment in the shared matrix row
#define TQB_S_SH_LEN 3
#define TQB_S_PS_BASE 4   // first float of this segment in every per-sample matrix row
#define TQB_S_PS_LEN 5
#define TQB_S_G_BASE 6    // first float of this segment in every gradient row
#define TQB_S_G_LEN 7

// group record: [regpos0..3, n_ops, first_op, n_ladders, kind]  (regpos = tile-local bit of register bit j;
// n_ladders: TQB_T_LADDER ops in the whole segment, first group record only; kind 1: ladder group - only
// TQB_T_LADDER ops, no register bits, applied directly on the staged tile)
// op record (12 words = 48 bytes; words 0-3 are read as one 128-bit word, two ops ahead of use):
#define TQB_OP_W0 0     // bits 0-1 target register bit j, 2-5 type, 8-15 flags, 16-19 gradient kind, 20-21 jc
#define TQB_OP_SLOT 1   // float offset of this op's matrix in the staged shared / per-sample floats
#define TQB_OP_EMASK 2  // low 16: elements e (register-bit patterns) whose register-bit controls hold;
                        // high 16: DIAG parity of e over the parity bits that are register bits
#define TQB_OP_PM_T 3   // DIAG: parity mask over the non-register tile-local bits
#define TQB_OP_CM_T 4   // control mask over the non-register tile-local bits
#define TQB_OP_CV_T 5   // required value of those bits
#define TQB_OP_CM_O 6   // control mask over the compressed outer (non-tile) index
#define TQB_OP_CV_O 7
#define TQB_OP_PM_O 8   // DIAG: parity mask over the compressed outer index
#define TQB_OP_GSLOT 9  // float offset of this op's gradient accumulators in the segment's gradient floats
#define TQB_OP_LPH 10   // TQB_T_LADDER: words 10-11 are SCRATCH in the staged copy: the kernel writes the tile's outer-bit phase
                        // (re, im as float bits) here once per tile, every thread reads it back
#define TQB_OP_PC(w0) (((w0) >> 22) & 15)   // TQB_F_PREH: tile-local position of the control bit

// op types.  Matrix slots are padded to a multiple of 4 floats (16-byte staged loads).
#define TQB_T_GEN 0    // 2x2 complex on register bit j; 8 floats row-major (re,im)
#define TQB_T_REAL 1   // real 2x2 on register bit j; 4 floats m00 m01 m10 m11       (ry, h, ...)
#define TQB_T_RX 2     // [[c, -is], [-is, c]] on register bit j; floats c, s, 0, 0     (rx)
#define TQB_T_DIAG 3   // amp *= d[parity]; 4 floats d0 (re,im), d1 (re,im)             (rz, cz, cu1, rzz ...)
#define TQB_T_PERM 4   // swap the two amplitudes of register bit j; no floats           (x, cnot, toffoli ...)
#define TQB_T_U2Q 5    // 4x4 complex on register bits (1,0) = (msb,lsb); 32 floats
// "Scaled" forms (programs that never run backward): the gate's common factor is left out of the per-amplitude
// arithmetic (one FMA per float instead of two FP instructions) and multiplied into the pass's deferred scale factor
// (TQB_H_SCALE_*), which is applied once when the tile is written back.  4 floats (c, s, r, mode):
#define TQB_T_RXS 6    // rx:  mode 0 (|c| >= |s|): c [[1, -ir], [-ir, 1]], r = s/c;  mode 1: -is [[ir, 1], [1, ir]], r = c/s
#define TQB_T_RYS 7    // ry:  mode 0: c [[1, -r], [r, 1]], r = s/c;                  mode 1: s [[r, -1], [1, r]], r = c/s
#define TQB_T_HS 8     // hadamard: a' = a + b, b' = a - b, factor 1/sqrt(2); no floats
#define TQB_T_DIAGS 9  // one-bit diagonal d0 diag(1, rel): amplitudes whose bit is set *= rel, factor d0; floats rel (re,im), d0 (re,im)
#define TQB_T_LADDER 10 // controlled phase ladder (a run of cu1 / cphase gates sharing one qubit, e.g. one rung set of a QFT):
                        // where the controls hold, amp *= prod over the ladder bits that are set of that bit's phase.
                        // floats: one (re, im) per tile-local bit position [T], then one per compressed outer bit;
                        // PM_T / PM_O: the ladder's tile-local / outer bits; controls in CM_T / CV_T / CM_O / CV_O
                        // (ladders run in groups of their own, without register bits)

#define TQB_F_TCTRL 0x100      // has controls on non-register tile bits (CM_T/CV_T valid)
#define TQB_F_RCTRL 0x200      // has controls on register bits (EMASK != 0xffff)
#define TQB_F_PS 0x400         // matrix lives in the per-sample floats
#define TQB_F_NOUNAPPLY 0x800  // backward: psi already holds the op's INPUT (non-unitary single-op node)
#define TQB_F_OCTRL 0x1000     // has controls on outer (non-tile) bits (CM_O/CV_O valid)
#define TQB_F_DJ 0x2000        // DIAG whose only register-resident parity bit is register bit j (element parity = bit j of e)
#define TQB_F_PREH 0x8000      // TQB_T_LADDER: an (unnormalised) Hadamard on the ladder's control bit is applied first, i.e. one rung set
                               // H_c; cu1(c+1, c); cu1(c+2, c) ... of a QFT in a single sweep over the tile; w0 bits 22-25: tile-local
                               // position of c; the 1/sqrt(2) joins the pass's deferred scale factor
#define TQB_F_POSTX 0x4000     // a CNOT (control = register bit JC, target = this op's bit j) is folded in: the two outputs
                               // of every pair whose control bit is set are exchanged (costs nothing: register renaming)

// gradient kinds (backward sweep): what is accumulated per sample for this op
#define TQB_G_NONE 0
#define TQB_G_FULL 1   // G = sum lam_out psi_in^H : 8 floats (GEN) / 32 floats (U2Q)
#define TQB_G_X 2      // w = sum Im(conj((X psi)_k) lam_k)  over active pairs: 1 float   (rx, crx)
#define TQB_G_Y 3      // w = sum Re(conj(psi_b) lam_a - conj(psi_a) lam_b):     1 float   (ry, cry)
#define TQB_G_DW 4     // w_p = sum_{parity p} Im(conj(psi_k) lam_k):            2 floats  (phases)
#define TQB_OP_TYPE(w0) (((w0) >> 2) & 15)
#define TQB_OP_J(w0) ((w0) & 3)
#define TQB_OP_CODE(type, j) (((type) << 2) | (j))
#define TQB_OP_GKIND(w0) (((w0) >> 16) & 15)
#define TQB_OP_JC(w0) (((w0) >> 20) & 3)   // TQB_F_POSTX: register bit of the folded CNOT's control

// matrix-builder table: one record of 10 words per op (tqb_build_mats / tqb_build_mats_bwd)
#define TQB_B_WORDS 10
#define TQB_B_KIND 0      // builder kind (below)
#define TQB_B_FLAGS 1     // bit0 inverse, bit1 per-sample params, bit2 matrix in per-sample row
#define TQB_B_POFF 2      // first parameter column (in params_sh or params_ps)
#define TQB_B_MOFF 3      // float offset in the matrix row (shared or per-sample)
#define TQB_B_COFF 4      // K_CONST: float offset into the constant pool
#define TQB_B_NFLOAT 5    // floats written
#define TQB_B_NPARAM 6
#define TQB_B_TYPE 7      // op type the matrix is packed for (TQB_T_*)
#define TQB_B_GKIND 8     // gradient kind accumulated by the pass kernel
#define TQB_B_GOFF 9      // float offset in the gradient row

#define TQB_BF_INVERSE 1
#define TQB_BF_PER_SAMPLE 2
// K_SCALE records: product of the scale factors of the scaled ops among table records [COFF, COFF + NFLOAT) -> 2 floats at MOFF

enum {
  TQB_K_CONST = 0,   // copy NFLOAT floats from the constant pool (inverse already folded in)
  TQB_K_EXTERN = 1,  // matrix written by the host front end (explicit unitary); builder skips it
  TQB_K_RX = 2, TQB_K_RY = 3, TQB_K_ROT = 4, TQB_K_U2 = 5, TQB_K_U2RAW = 6, TQB_K_U3 = 7,
  TQB_K_CU = 8, TQB_K_R = 9,                     // -> 2x2
  TQB_K_RZ = 10, TQB_K_PHASE = 11, TQB_K_GPHASE = 12,  // -> diagonal pair d0,d1
  TQB_K_RXX = 13, TQB_K_RYY = 14, TQB_K_RZX = 15, TQB_K_SEXC = 16, TQB_K_XXMYY = 17,
  TQB_K_XXPYY = 18,                               // -> 4x4
  TQB_K_SCALE = 19,  // deferred scale factor of one pass (see TQB_BF_* above)
  TQB_K_COUNT = 20
};

// ---------------------------------------------------------------------------------------
// Block programs (forward-only, streamed states): runs of 1-2(-6)-qubit gates multiplied on the device into dense
// 6-qubit block unitaries (tqb_block_build), each applied in one tensor-core pass (tqb_apply_blocks).
// Block header: 12 words per block.
#define TQB_KB_WORDS 12
#define TQB_KB_OPFIRST 0   // first op record of the block
#define TQB_KB_NOPS 1
#define TQB_KB_FLAGS 2     // TQB_KBF_*
#define TQB_KB_KPOS 4      // 6 words: physical bit positions (distinct, all >= 6, any order) of block-local bits 0..5
#define TQB_KBF_TRANSPOSE 1  // transposed write-back: block-local bits and index bits 0-5 exchange their roles on the way out
#define TQB_KBF_ROTATE 2     // out of place: the tile is written contiguously (block bits -> index bits 6-11, the other
                             // bits >= 6 compacted above them in ascending order); see tqb_apply_blocks
// Block op record: 8 words.  Bits are block-local (0..5) unless stated.
#define TQB_KO_WORDS 8
#define TQB_KO_W0 0        // bits 0-3 form, 4-7 number of parameters, bit 8 inverse
#define TQB_KO_BITS 1      // bits 0-7 target t0 (matrix lsb), 8-15 target t1 (matrix msb, 2-qubit forms)
#define TQB_KO_MASKS 2     // bits 0-7 control mask, 8-15 control value, 16-23 parity mask (diagonal forms)
#define TQB_KO_RCM 3       // sharded states: control mask over the rank index ...
#define TQB_KO_RCV 4       // ... its required value ...
#define TQB_KO_RPM 5       // ... and the parity mask over the rank index (diagonal forms)
#define TQB_KO_KIND 6      // TQB_K_CONST or a parametric builder kind (TQB_K_RX ...)
#define TQB_KO_OFF 7       // K_CONST: float offset in the constant pool (full complex form, inverse folded in);
                           // parametric: first parameter (float offset in params_sh)
// Quadratic-phase diagonal passes (tqb_quadphase): coefficient words, uint32 fractions of a turn:
// [0] c0, [1 + b] alpha_b, [1 + TQB_QP_BITS + a * TQB_QP_BITS + b] beta_ab (a < b) over physical index bits b (rank bits above the local ones)
#define TQB_QP_BITS 40
#define TQB_QP_WORDS 1641
#define TQB_KF_U1 0        // 2x2 complex on t0 where the controls hold            (consts: 8 floats)
#define TQB_KF_DIAG 1      // amp *= d[parity] where the controls hold               (consts: 4 floats)
#define TQB_KF_U2 2        // 4x4 complex on (t1, t0) where the controls hold      (consts: 32 floats)
#define TQB_KF_X 3         // exchange the pair of t0 where the controls hold        (no floats)
