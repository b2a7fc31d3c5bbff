/* libtqb200 — C ABI of the B200-native statevector engine that drops in behind TorchQuantum's
 * gate-application hot path.  Plain pointers and sizes only; every pointer is DEVICE memory
 * unless its name ends in _host; `stream` is a cudaStream_t passed as void*.  All functions
 * return 0 on success, -1 on failure (message via tqb_last_error()).  Nothing here falls back
 * to the CPU.
 *
 * State layout: complex64 interleaved (re, im) float pairs, [bsz][2^n], wire w of the reference
 * (torchquantum/device/devices.py:57-71: dim 1+w of the [bsz,2,...,2] tensor) is index bit
 * n-1-w.  `prog_dev` is a gate program in the wire format of torchquantum_b200/csrc/tqb_program.h
 * produced by the host planner (torchquantum_b200/plan.py).
 *
 * Reference interfaces replaced (paths relative to the reference repo's torchquantum/):
 *   tqb_build_mats        <- the 26 `*_matrix(params)` builders + constant tables
 *                            (functional/rx.py:19-115, ry.py:19-104, rz.py:19-171, u3.py:19-112,
 *                            rot.py:19-89, u1.py, u2.py, r.py, phase_shift.py, single_excitation.py,
 *                            xx_min_yy.py, xx_plus_yy.py, global_phase.py; mat_dict
 *                            functional/__init__.py:97-123) incl. `inverse` (gate_wrapper.py:427-432)
 *   tqb_pass_forward      <- apply_unitary_bmm / apply_unitary_einsum, one call per gate
 *                            (functional/gate_wrapper.py:96-140, :18-93), here many gates per pass
 *   tqb_pass_backward +   <- torch autograd through every bmm/permute/cos/sin node of the above
 *   tqb_build_mats_bwd       (exercised by examples/mnist/mnist.py:143-146); here an adjoint sweep
 *   tqb_apply_dense(_tc)  <- the same bmm for arbitrary 2^k x 2^k matrices (qft, qubitunitary*)
 *   tqb_expval_pauli      <- expval_joint_analytical (measurement/measurements.py:232-279)
 *   tqb_expval_z          <- the marginal/eigenvalue part of expval / MeasureAll (:297-312)
 *   tqb_probs, tqb_sample_counts <- measure (:37-74): |psi|^2 on the host + random.choices
 */
#ifndef TQB200_H
#define TQB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQB_ABI_VERSION 7

int tqb_abi_version(void);
const char* tqb_last_error(void);

/* Gate matrices live in one row shared by the batch (`msh`: constants and gates with shared
 * parameters) and one row per sample (`mps[bsz][mps_stride]`: per-sample parameters).
 * table_dev: n_records x 10 int32 (csrc/tqb_program.h, TQB_B_*); the first n_shared_records write
 * `msh`, the others write every sample's `mps` row.  params_sh: [P_sh] shared parameters;
 * params_ps: [bsz][ps_stride] per-sample parameters. */
int tqb_build_mats(const int32_t* table_dev, int n_records, int n_shared_records, long long bsz,
                   const float* params_sh, const float* params_ps, int ps_stride, const float* consts,
                   float* msh, float* mps, long long mps_stride, void* stream);

/* dL/dtheta from the per-sample gradient rows gr[bsz][gr_stride] written by tqb_pass_backward:
 * per-sample parameters -> gparams_ps[bsz][ps_stride]; shared parameters -> per-sample partials
 * gparams_sh_partial[bsz][sh_stride] (caller sums over bsz). */
int tqb_build_mats_bwd(const int32_t* table_dev, int n_records, long long bsz, const float* params_sh,
                       const float* params_ps, int ps_stride, const float* gr, long long gr_stride,
                       float* gparams_sh_partial, int sh_stride, float* gparams_ps, void* stream);

/* One fused pass over the state: every tile of 2^T amplitudes is read from `src`, staged on chip,
 * the whole gate list of the pass applied to it, and written to `dst` (dst may equal src).
 * desc_host: 8 ints on the HOST {T, R, L, n_seg, sh_max, ps_max, g_max, 0} copied from the program
 * header (tile geometry and staging capacities).
 * ext_hi: bits OR-ed above the tile index when gates are controlled by / diagonal in qubits that
 * live on other ranks (sharded statevector). */
int tqb_pass_forward(const float* src, float* dst, int n_qubits, long long bsz, const int32_t* prog_dev,
                     const int* desc_host, const float* msh, const float* mps, long long mps_stride,
                     unsigned ext_hi, void* stream);

/* Adjoint sweep of the same pass: reads the pass OUTPUT psi and dL/dpsi_out, writes the pass INPUT
 * psi and dL/dpsi_in (dst may alias src); writes (single tile per sample) or atomically
 * accumulates (several tiles: pre-zero) the per-sample gradient rows gr[bsz][gr_stride]. */
int tqb_pass_backward(const float* psi_src, const float* lam_src, float* psi_dst, float* lam_dst, int n_qubits,
                      long long bsz, const int32_t* prog_dev, const int* desc_host, const float* msh,
                      const float* mps, long long mps_stride, float* gr, long long gr_stride, unsigned ext_hi,
                      void* stream);

/* out = U_wires in (out of place); mats: complex64 [2^k][2^k] per sample with stride
 * mat_stride_complex (0 = shared); bitpos_dev: k bit positions, matrix MSB first;
 * conj_transpose != 0 applies U^H. */
int tqb_apply_dense(const float* in, float* out, int n_qubits, long long bsz, const float* mats,
                    long long mat_stride_complex, int k, const int* bitpos_dev, int conj_transpose,
                    void* stream);
/* Tensor-core form of tqb_apply_dense for ONE matrix shared by the batch, 4 <= k <= 7 and n_qubits >= 12
 * (mma.sync TF32 with a 3xTF32 error-compensated split).  bitpos_host: k bit positions on the
 * HOST, matrix MSB first.  Replaces the same reference code as tqb_apply_dense (functional/qft.py:19-33,
 * functional/qubit_unitary.py:19-93 through gate_wrapper.py:96-140). */
int tqb_apply_dense_tc(const float* in, float* out, int n_qubits, long long bsz, const float* mat, int k,
                       const int* bitpos_host, int conj_transpose, void* stream);
/* Forward-only block programs (csrc/tqb_program.h "Block programs"): runs of gates on <= 6 qubits are multiplied on the
 * device into dense 6-qubit block unitaries, each applied in ONE tensor-core pass over the state (tcgen05.mma kind::tf32
 * with the 3xTF32 split, TMEM accumulators, TMA tile loads / stores).  Replaces what the reference's static mode does on
 * the host (graph/graphs.py:384-509 block scheduling, :713-861 block unitaries, one bmm per block through
 * functional/gate_wrapper.py:96-140).
 * tqb_block_build: blocks_dev [n_blocks][TQB_KB_WORDS], ops_dev [*][TQB_KO_WORDS], constant pool, shared parameters
 *   -> a_planes [n_blocks][2][128][64] float (TF32 hi / lo planes of A[2i+w][j] = (Re, Im)[w] U[i][j]).  `rank` resolves
 *   controls / parity bits that live in the rank index of a sharded state.
 * tqb_apply_blocks: applies the blocks in order; the first reads src and the result ends up in dst.  A block without
 *   TQB_KBF_ROTATE writes its tiles back where they came from (in place once the state is in dst / scratch).  A block
 *   with TQB_KBF_ROTATE reads its tiles from wherever its six bits are and writes each as ONE contiguous 32 KiB piece
 *   into the other buffer of {dst, scratch}: afterwards the block bits sit in index bits 6-11 (pattern order) and all
 *   other bits >= 6 follow above them in their previous relative order (gathering reads run at full speed for every
 *   placement, scattered 512-byte writes at strides >= 4 MiB do not; the caller tracks the qubit layout).  scratch may
 *   be null when no block rotates; with an odd number of rotating blocks dst must differ from src (src may then serve
 *   as scratch).  blocks_host: the same block headers on the HOST (bit positions for the tensor maps).  n_qubits >= 12.
 *   renorm != 0: every block is unitary, tiles are rescaled to the norm they came in with. */
int tqb_block_build(const int32_t* blocks_dev, int n_blocks, const int32_t* ops_dev, const float* consts,
                    const float* params_sh, unsigned rank, float* a_planes, void* stream);
int tqb_apply_blocks(const float* src, float* dst, float* scratch, int n_qubits, long long bsz, const float* a_planes,
                     const int32_t* blocks_host, int n_blocks, int renorm, void* stream);
/* Fused quadratic-phase diagonal pass: state[b][x] *= exp(2 pi i (c0 + sum_k alpha_k x_k + sum_{j<k} beta_jk x_j x_k)) with
 * x_k = bit k of the amplitude index (bits >= n_local: the bits of `rank`, sharded states), in place.  coef_dev:
 * TQB_QP_WORDS uint32 fractions of a turn (csrc/tqb_program.h).  Any number of commuting 1- / 2-qubit diagonal gates
 * (rz, phaseshift, cz, cu1, crz, rzz, folded cnot-rz-cnot sandwiches; constants or shared parameters) in ONE trip
 * through HBM; replaces one apply_unitary_bmm per gate (functional/gate_wrapper.py:96-140 via functional/rz.py,
 * functional/phase_shift.py, functional/u1.py ...).  n_local >= 12, n_local <= n_bits <= TQB_QP_BITS. */
int tqb_quadphase(float* state, int n_local, long long bsz, const uint32_t* coef_dev, int n_bits, unsigned rank, void* stream);
/* G[b][r][c] = sum_rest lam_out[b][r,rest] conj(psi_in[b][c,rest]) (complex64 [bsz][2^k][2^k]) */
int tqb_dense_grad(const float* psi_in, const float* lam_out, int n_qubits, long long bsz, int k,
                   const int* bitpos_dev, float* G, void* stream);

/* out[bsz][n_terms] = <psi|P_t|psi>, P_t given by x/z bit masks and the number of Y factors.
 * Pre-zero `out` when n_qubits > 12 (several blocks per sample accumulate atomically). */
int tqb_expval_pauli(const float* state, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                     const unsigned long long* zmask_dev, const int* ny_dev, int n_terms, float* out,
                     void* stream);
int tqb_expval_pauli_bwd(const float* state, int n_qubits, long long bsz, const unsigned long long* xmask_dev,
                         const unsigned long long* zmask_dev, const int* ny_dev, int n_terms,
                         const float* grad_out, float* grad_state, void* stream);

/* out[bsz][n+1]: column j < n is <Z> of index bit j, column n is the squared norm.
 * Pre-zero `out` when n_qubits > 12. */
int tqb_expval_z(const float* state, int n_qubits, long long bsz, float* out, void* stream);
int tqb_expval_z_bwd(const float* state, int n_qubits, long long bsz, const float* grad_out, float* grad_state,
                     void* stream);

/* probs[bsz][2^n] float32 with the reference's host rounding (float32 hypot, then squared). */
int tqb_probs(const float* state, int n_qubits, long long bsz, float* probs, void* stream);
/* counts[bsz][2^n] (pre-zeroed int32) += histogram of n_shots inverse-CDF draws per sample from
 * uniforms[bsz][n_shots] (float32 of Python's random.random()), CDF = sequential float32 sums. */
int tqb_sample_counts(const float* probs, int n_qubits, long long bsz, const float* uniforms, int n_shots,
                      float* cdf_scratch, int* counts, void* stream);

/* Batch collapse (engine.py): a run of gates shared by the whole batch is one 2^n x 2^n unitary, applied to all states at
 * once as a GEMM on the tensor cores; what the reference's static mode does block-wise with torch.bmm
 * (graph/graphs.py:713-861 block unitaries, :836-861 QuantumGraph.apply_unitary).
 * tqb_split_tf32: x[n] -> TF32 hi and lo parts (n % 4 == 0).
 * tqb_gemm_tf32x3: C[M][N] = A[M][K] * Bt[N][K]^T (all row-major fp32, operands given as their hi / lo parts) with fp32
 *   accuracy: Alo Bhi + Ahi Blo + Ahi Bhi, tcgen05.mma kind::tf32.  M % 128 == N % 256 == K % 32 == 0.  K is cut into chains
 *   of <= 512 (the tensor core's truncating accumulation is biased in proportion to the chain length); the partial
 *   products go to `workspace` (tqb_gemm_workspace_floats(M, N, K) floats, 0 = not needed) and are summed in fp32
 *   round-to-nearest.  A complex product is this GEMM on the interleaved (re, im) views with Bt the real 2x2 embedding of
 *   the right factor. */
int tqb_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);
/* Operands of the complex GEMM from a complex64 matrix x [rows][cols] (row-major, interleaved re / im) in one pass,
 * already cut into TF32 hi / lo planes (rows, cols multiples of 32):
 *   mode 1: A of x^H:  out [cols][2 rows],   out[c][2r + v]      = (re, -im)[v] x[r][c]
 *   mode 2: Bt of x:   out [2 cols][2 rows], out[2c][2r + v]     = (re, -im)[v] x[r][c],  out[2c + 1][2r + v] = (im, re)[v] x[r][c]
 * so that x^H . g = tqb_gemm_tf32x3(A(mode 1 of x), Bt(mode 2 of g)) and a . b = tqb_gemm_tf32x3(split(a), Bt(mode 2 of b)). */
int tqb_gemm_prep(const float* x, long long rows, long long cols, int mode, float* hi, float* lo, void* stream);
long long tqb_gemm_workspace_floats(long long M, long long N, long long K);
int tqb_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* bt_hi, const float* bt_lo, float* c, float* workspace,
                    long long M, long long N, long long K, void* stream);

/* Sharded statevector (no counterpart in the reference, which is single-device: device/devices.py:36-136; README.md:52 only
 * claims multi-GPU): swap the g rank-index qubits with the g local index bits victim_bits_host[0..g) (ascending) by
 * writing every amplitude of `src` (this rank's slab [bsz][2^n_local]) straight into its place in the destination slab of
 * the rank that owns it afterwards.  peer_dst_host: 2^g device pointers (HOST array), entry r = destination slab of rank r
 * mapped into this process (CUDA IPC / symmetric memory), entry `rank` = the local destination (must differ from src).
 * The caller synchronises the ranks before the call (nobody is still using a destination slab) and after it (all slabs
 * complete) before anybody reads a destination.  Between two ranks, exchanges whose lowest victim bit leaves contiguous
 * runs of >= 8 MiB are issued as peer copies on the copy engines instead of the store kernel (TQB200_EXCHANGE_CE = 0 / 1 /
 * 2: never / two ranks only / any number of ranks). */
int tqb_exchange_p2p(const float* src, float* const* peer_dst_host, int n_local, long long bsz, int g,
                     const int* victim_bits_host, int rank, void* stream);

/* HOST function (no device work): out_host[i] = float32(random.random()) for the next n draws of CPython's Mersenne
 * Twister, whose state (`random.getstate()[1]`: 624 words + position) is advanced in place.  Replaces the interpreter loop
 * behind `random.choices(..., k=n_shots)` in measure (measurement/measurements.py:37-74); the stream itself must stay the
 * host's for histograms to be bit-exact with the reference. */
int tqb_host_uniforms(uint32_t* mt_state625_host, long long n, float* out_host);

#ifdef __cplusplus
}
#endif
#endif
