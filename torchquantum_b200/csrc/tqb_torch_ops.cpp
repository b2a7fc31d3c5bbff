// torch.library registration of the libtqb200 C ABI (include/tqb200.h): namespace `tqb200`, one custom op per ABI entry
// point, same argument order; device pointers become Tensors, host arrays become int[] / CPU tensors, the stream is
// the current CUDA stream of the first tensor's device.  This file is the whole "thin C++ extension" north_star asks
// for: it holds no arithmetic and no torch types cross into libtqb200.so, which stays a plain C ABI resolved with
// dlopen/dlsym (so the reference-side stub in INTEGRATION.md can bind the same library without torch).
//
// What sits behind these ops in the reference: torchquantum/functional/gate_wrapper.py:324-446 (gate_wrapper ->
// apply_unitary_bmm :96-140), the matrix builders of functional/*.py, torchquantum/measurement/measurements.py:37-343.
//
// Dispatch keys: CUDA -> the sm_100a library.  CPU -> ONLY the host emulation build that the test-suite injects with
// tqb200::_bind(path, emulated=True); without it a CPU tensor raises (there is no CPU fallback).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <dlfcn.h>
#include <torch/library.h>

#include <string>

#include "../../include/tqb200.h"
#include "tqb_program.h"

namespace {

struct Abi {
  void* handle = nullptr;
  decltype(&tqb_abi_version) abi_version = nullptr;
  decltype(&tqb_last_error) last_error = nullptr;
  decltype(&tqb_build_mats) build_mats = nullptr;
  decltype(&tqb_build_mats_bwd) build_mats_bwd = nullptr;
  decltype(&tqb_pass_forward) pass_forward = nullptr;
  decltype(&tqb_pass_backward) pass_backward = nullptr;
  decltype(&tqb_apply_dense) apply_dense = nullptr;
  decltype(&tqb_apply_dense_tc) apply_dense_tc = nullptr;
  decltype(&tqb_block_build) block_build = nullptr;
  decltype(&tqb_apply_blocks) apply_blocks = nullptr;
  decltype(&tqb_dense_grad) dense_grad = nullptr;
  decltype(&tqb_quadphase) quadphase = nullptr;
  decltype(&tqb_expval_pauli) expval_pauli = nullptr;
  decltype(&tqb_expval_pauli_bwd) expval_pauli_bwd = nullptr;
  decltype(&tqb_expval_z) expval_z = nullptr;
  decltype(&tqb_expval_z_bwd) expval_z_bwd = nullptr;
  decltype(&tqb_probs) probs = nullptr;
  decltype(&tqb_sample_counts) sample_counts = nullptr;
  decltype(&tqb_host_uniforms) host_uniforms = nullptr;
  decltype(&tqb_exchange_p2p) exchange_p2p = nullptr;
  decltype(&tqb_split_tf32) split_tf32 = nullptr;
  decltype(&tqb_gemm_prep) gemm_prep = nullptr;
  decltype(&tqb_gemm_tf32x3) gemm_tf32x3 = nullptr;
  decltype(&tqb_gemm_workspace_floats) gemm_workspace_floats = nullptr;
};

Abi g_cuda, g_emu;

template <class F>
void resolve(void* h, const char* name, F& fn, bool required) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  TORCH_CHECK(fn != nullptr || !required, "libtqb200: symbol ", name, " not found");
}

int64_t bind(const std::string& path, bool emulated) {
  Abi& a = emulated ? g_emu : g_cuda;
  if (path.empty()) {  // unbind (tests switch back to the CUDA library)
    a = Abi();
    return 0;
  }
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  TORCH_CHECK(h != nullptr, "libtqb200: cannot load ", path, ": ", dlerror());
  Abi n;
  n.handle = h;
  resolve(h, "tqb_abi_version", n.abi_version, true);
  resolve(h, "tqb_last_error", n.last_error, true);
  resolve(h, "tqb_build_mats", n.build_mats, true);
  resolve(h, "tqb_build_mats_bwd", n.build_mats_bwd, true);
  resolve(h, "tqb_pass_forward", n.pass_forward, true);
  resolve(h, "tqb_pass_backward", n.pass_backward, true);
  resolve(h, "tqb_apply_dense", n.apply_dense, true);
  resolve(h, "tqb_apply_dense_tc", n.apply_dense_tc, !emulated);
  resolve(h, "tqb_block_build", n.block_build, true);
  resolve(h, "tqb_apply_blocks", n.apply_blocks, true);
  resolve(h, "tqb_dense_grad", n.dense_grad, true);
  resolve(h, "tqb_quadphase", n.quadphase, true);
  resolve(h, "tqb_expval_pauli", n.expval_pauli, true);
  resolve(h, "tqb_expval_pauli_bwd", n.expval_pauli_bwd, true);
  resolve(h, "tqb_expval_z", n.expval_z, true);
  resolve(h, "tqb_expval_z_bwd", n.expval_z_bwd, true);
  resolve(h, "tqb_probs", n.probs, true);
  resolve(h, "tqb_sample_counts", n.sample_counts, true);
  resolve(h, "tqb_host_uniforms", n.host_uniforms, true);
  resolve(h, "tqb_exchange_p2p", n.exchange_p2p, true);
  resolve(h, "tqb_split_tf32", n.split_tf32, true);
  resolve(h, "tqb_gemm_prep", n.gemm_prep, true);
  resolve(h, "tqb_gemm_tf32x3", n.gemm_tf32x3, true);
  resolve(h, "tqb_gemm_workspace_floats", n.gemm_workspace_floats, true);
  TORCH_CHECK(n.abi_version() == TQB_ABI_VERSION, "libtqb200: ABI version ", n.abi_version(), " != ", TQB_ABI_VERSION);
  a = n;
  return n.abi_version();
}

const Abi& abi_for(const at::Tensor& t) {
  if (t.is_cuda()) {
    TORCH_CHECK(g_cuda.handle != nullptr, "torchquantum_b200: the CUDA library libtqb200.so is not loaded (run "
                                          "__graft_entry__.build()); there is no CPU fallback");
    return g_cuda;
  }
  TORCH_CHECK(g_emu.handle != nullptr,
              "torchquantum_b200 runs on CUDA (sm_100a) only: got a CPU tensor and there is no CPU fallback");
  return g_emu;
}

// The kernels launch on the CURRENT device: make it the tensor's device for the duration of the op (a state on cuda:1
// while cuda:0 is current would otherwise launch on the wrong device with a foreign stream).
#define TQB_ABI(a, t)                            \
  c10::cuda::OptionalCUDAGuard tqb_device_guard; \
  if ((t).is_cuda()) tqb_device_guard.set_device((t).device()); \
  const Abi& a = abi_for(t)

void* stream_of(const at::Tensor& t) {
  return t.is_cuda() ? static_cast<void*>(c10::cuda::getCurrentCUDAStream(t.device().index()).stream()) : nullptr;
}

void check(const Abi& a, int rc, const char* what) { TORCH_CHECK(rc == 0, what, " failed: ", a.last_error()); }

const float* fp(const at::Tensor& t) {
  TORCH_CHECK(t.is_contiguous(), "libtqb200 ops take contiguous tensors");
  return static_cast<const float*>(t.data_ptr());
}
const float* fp(const c10::optional<at::Tensor>& t) { return t.has_value() && t->defined() ? fp(*t) : nullptr; }
float* fpm(at::Tensor& t) {
  TORCH_CHECK(t.is_contiguous(), "libtqb200 ops take contiguous tensors");
  return static_cast<float*>(t.data_ptr());
}
const int32_t* ip(const at::Tensor& t) {
  TORCH_CHECK(t.is_contiguous() && t.scalar_type() == at::kInt, "libtqb200: expected a contiguous int32 tensor");
  return static_cast<const int32_t*>(t.data_ptr());
}
const unsigned long long* up(const at::Tensor& t) {
  TORCH_CHECK(t.is_contiguous() && t.element_size() == 8, "libtqb200: expected a contiguous 64-bit mask tensor");
  return static_cast<const unsigned long long*>(t.data_ptr());
}

void build_mats(const at::Tensor& table, int64_t n_records, int64_t n_shared, int64_t bsz,
                const c10::optional<at::Tensor>& params_sh, const c10::optional<at::Tensor>& params_ps, int64_t ps_stride,
                const c10::optional<at::Tensor>& consts, at::Tensor& msh, at::Tensor& mps, int64_t mps_stride) {
  TQB_ABI(a, msh);
  check(a, a.build_mats(ip(table), (int)n_records, (int)n_shared, bsz, fp(params_sh), fp(params_ps), (int)ps_stride, fp(consts),
                        fpm(msh), fpm(mps), mps_stride, stream_of(msh)), "tqb_build_mats");
}

void build_mats_bwd(const at::Tensor& table, int64_t n_records, int64_t bsz, const c10::optional<at::Tensor>& params_sh,
                    const c10::optional<at::Tensor>& params_ps, int64_t ps_stride, const at::Tensor& gr, int64_t gr_stride,
                    at::Tensor& gsh_partial, int64_t sh_stride, at::Tensor& gps) {
  TQB_ABI(a, gr);
  check(a, a.build_mats_bwd(ip(table), (int)n_records, bsz, fp(params_sh), fp(params_ps), (int)ps_stride, fp(gr), gr_stride,
                            fpm(gsh_partial), (int)sh_stride, fpm(gps), stream_of(gr)), "tqb_build_mats_bwd");
}

void pass_forward(const at::Tensor& src, at::Tensor& dst, int64_t n, int64_t bsz, const at::Tensor& prog, at::IntArrayRef desc,
                  const at::Tensor& msh, const at::Tensor& mps, int64_t mps_stride, int64_t ext_hi) {
  TQB_ABI(a, src);
  TORCH_CHECK(desc.size() == 8, "tqb_pass_forward: desc holds 8 ints");
  int d[8];
  for (int i = 0; i < 8; ++i) d[i] = (int)desc[i];
  check(a, a.pass_forward(fp(src), fpm(dst), (int)n, bsz, ip(prog), d, fp(msh), fp(mps), mps_stride, (unsigned)ext_hi,
                          stream_of(src)), "tqb_pass_forward");
}

void pass_backward(const at::Tensor& psi_src, const at::Tensor& lam_src, at::Tensor& psi_dst, at::Tensor& lam_dst, int64_t n,
                   int64_t bsz, const at::Tensor& prog, at::IntArrayRef desc, const at::Tensor& msh, const at::Tensor& mps,
                   int64_t mps_stride, at::Tensor& gr, int64_t gr_stride, int64_t ext_hi) {
  TQB_ABI(a, psi_src);
  TORCH_CHECK(desc.size() == 8, "tqb_pass_backward: desc holds 8 ints");
  int d[8];
  for (int i = 0; i < 8; ++i) d[i] = (int)desc[i];
  check(a, a.pass_backward(fp(psi_src), fp(lam_src), fpm(psi_dst), fpm(lam_dst), (int)n, bsz, ip(prog), d, fp(msh), fp(mps),
                           mps_stride, fpm(gr), gr_stride, (unsigned)ext_hi, stream_of(psi_src)), "tqb_pass_backward");
}

void apply_dense(const at::Tensor& in, at::Tensor& out, int64_t n, int64_t bsz, const at::Tensor& mats, int64_t mat_stride,
                 int64_t k, const at::Tensor& bitpos, int64_t conj_t) {
  TQB_ABI(a, in);
  check(a, a.apply_dense(fp(in), fpm(out), (int)n, bsz, fp(mats), mat_stride, (int)k, ip(bitpos), (int)conj_t, stream_of(in)),
        "tqb_apply_dense");
}

void apply_dense_tc(const at::Tensor& in, at::Tensor& out, int64_t n, int64_t bsz, const at::Tensor& mat, int64_t k,
                    at::IntArrayRef bitpos_host, int64_t conj_t) {
  TQB_ABI(a, in);
  TORCH_CHECK(a.apply_dense_tc != nullptr, "tqb_apply_dense_tc: not available in this build");
  TORCH_CHECK((int64_t)bitpos_host.size() == k && k <= 16, "tqb_apply_dense_tc: bitpos holds k ints");
  int b[16];
  for (int64_t i = 0; i < k; ++i) b[i] = (int)bitpos_host[i];
  check(a, a.apply_dense_tc(fp(in), fpm(out), (int)n, bsz, fp(mat), (int)k, b, (int)conj_t, stream_of(in)), "tqb_apply_dense_tc");
}

void block_build(const at::Tensor& blocks, int64_t n_blocks, const at::Tensor& ops, const at::Tensor& consts,
                 const at::Tensor& params_sh, int64_t rank, at::Tensor& a_planes) {
  TQB_ABI(a, a_planes);
  check(a, a.block_build(ip(blocks), (int)n_blocks, ip(ops), fp(consts), fp(params_sh), (unsigned)rank, fpm(a_planes),
                         stream_of(a_planes)), "tqb_block_build");
}

void apply_blocks(const at::Tensor& src, at::Tensor& dst, const c10::optional<at::Tensor>& scratch, int64_t n, int64_t bsz,
                  const at::Tensor& a_planes, const at::Tensor& blocks_host, int64_t n_blocks, int64_t renorm) {
  TQB_ABI(a, src);
  TORCH_CHECK(!blocks_host.is_cuda(), "tqb_apply_blocks: the block headers are read on the host");
  float* sc = nullptr;
  if (scratch.has_value() && scratch->defined()) {
    TORCH_CHECK(scratch->is_contiguous() && scratch->device() == src.device() && scratch->nbytes() >= dst.nbytes(),
                "tqb_apply_blocks: scratch must be a contiguous state of the same size on the same device");
    sc = reinterpret_cast<float*>(scratch->data_ptr());
  }
  check(a, a.apply_blocks(fp(src), fpm(dst), sc, (int)n, bsz, fp(a_planes), ip(blocks_host), (int)n_blocks, (int)renorm,
                          stream_of(src)), "tqb_apply_blocks");
}

void quadphase(at::Tensor& state, int64_t n_local, int64_t bsz, const at::Tensor& coef, int64_t n_bits, int64_t rank) {
  TQB_ABI(a, state);
  TORCH_CHECK(coef.is_contiguous() && coef.scalar_type() == at::kInt && coef.numel() >= TQB_QP_WORDS && coef.device() == state.device(),
              "tqb_quadphase: coef must be TQB_QP_WORDS int32 words on the state's device");
  check(a, a.quadphase(fpm(state), (int)n_local, bsz, reinterpret_cast<const uint32_t*>(coef.data_ptr()), (int)n_bits,
                       (unsigned)rank, stream_of(state)), "tqb_quadphase");
}

void dense_grad(const at::Tensor& psi_in, const at::Tensor& lam_out, int64_t n, int64_t bsz, int64_t k, const at::Tensor& bitpos,
                at::Tensor& G) {
  TQB_ABI(a, psi_in);
  check(a, a.dense_grad(fp(psi_in), fp(lam_out), (int)n, bsz, (int)k, ip(bitpos), fpm(G), stream_of(psi_in)), "tqb_dense_grad");
}

void expval_pauli(const at::Tensor& state, int64_t n, int64_t bsz, const at::Tensor& xm, const at::Tensor& zm, const at::Tensor& ny,
                  int64_t n_terms, at::Tensor& out) {
  TQB_ABI(a, state);
  check(a, a.expval_pauli(fp(state), (int)n, bsz, up(xm), up(zm), ip(ny), (int)n_terms, fpm(out), stream_of(state)),
        "tqb_expval_pauli");
}

void expval_pauli_bwd(const at::Tensor& state, int64_t n, int64_t bsz, const at::Tensor& xm, const at::Tensor& zm,
                      const at::Tensor& ny, int64_t n_terms, const at::Tensor& grad_out, at::Tensor& grad_state) {
  TQB_ABI(a, state);
  check(a, a.expval_pauli_bwd(fp(state), (int)n, bsz, up(xm), up(zm), ip(ny), (int)n_terms, fp(grad_out), fpm(grad_state),
                              stream_of(state)), "tqb_expval_pauli_bwd");
}

void expval_z(const at::Tensor& state, int64_t n, int64_t bsz, at::Tensor& out) {
  TQB_ABI(a, state);
  check(a, a.expval_z(fp(state), (int)n, bsz, fpm(out), stream_of(state)), "tqb_expval_z");
}

void expval_z_bwd(const at::Tensor& state, int64_t n, int64_t bsz, const at::Tensor& grad_out, at::Tensor& grad_state) {
  TQB_ABI(a, state);
  check(a, a.expval_z_bwd(fp(state), (int)n, bsz, fp(grad_out), fpm(grad_state), stream_of(state)), "tqb_expval_z_bwd");
}

void probs(const at::Tensor& state, int64_t n, int64_t bsz, at::Tensor& out) {
  TQB_ABI(a, state);
  check(a, a.probs(fp(state), (int)n, bsz, fpm(out), stream_of(state)), "tqb_probs");
}

void sample_counts(const at::Tensor& probs_t, int64_t n, int64_t bsz, const at::Tensor& uniforms, int64_t n_shots,
                   at::Tensor& cdf_scratch, at::Tensor& counts) {
  TQB_ABI(a, probs_t);
  TORCH_CHECK(counts.is_contiguous() && counts.scalar_type() == at::kInt, "tqb_sample_counts: counts is int32");
  check(a, a.sample_counts(fp(probs_t), (int)n, bsz, fp(uniforms), (int)n_shots, fpm(cdf_scratch),
                           static_cast<int*>(counts.data_ptr()), stream_of(probs_t)), "tqb_sample_counts");
}

// peer_ptrs: the destination slab of every rank as raw device addresses (symmetric memory: _SymmetricMemory.buffer_ptrs)
void exchange_p2p(const at::Tensor& src, at::Tensor& local_dst, at::IntArrayRef peer_ptrs, int64_t n_local, int64_t bsz,
                  at::IntArrayRef victim_bits, int64_t rank) {
  TQB_ABI(a, src);
  const int g = (int)victim_bits.size();
  TORCH_CHECK(g >= 1 && g <= 4 && (int64_t)peer_ptrs.size() == (1ll << g), "tqb_exchange_p2p: 2^g peer pointers for g victim bits");
  TORCH_CHECK(rank >= 0 && rank < (int64_t)peer_ptrs.size() &&
                  reinterpret_cast<void*>((uintptr_t)peer_ptrs[rank]) == local_dst.data_ptr(),
              "tqb_exchange_p2p: peer_ptrs[rank] must be the local destination slab");
  float* ptrs[16];
  int vb[4];
  for (size_t r = 0; r < peer_ptrs.size(); ++r) ptrs[r] = reinterpret_cast<float*>((uintptr_t)peer_ptrs[r]);
  for (int k = 0; k < g; ++k) vb[k] = (int)victim_bits[k];
  check(a, a.exchange_p2p(fp(src), ptrs, (int)n_local, bsz, g, vb, (int)rank, stream_of(src)), "tqb_exchange_p2p");
}

void split_tf32(const at::Tensor& x, at::Tensor& hi, at::Tensor& lo) {
  TQB_ABI(a, x);
  TORCH_CHECK(x.scalar_type() == at::kFloat && hi.numel() == x.numel() && lo.numel() == x.numel(), "tqb_split_tf32: float32, equal sizes");
  check(a, a.split_tf32(fp(x), fpm(hi), fpm(lo), x.numel(), stream_of(x)), "tqb_split_tf32");
}

void gemm_prep(const at::Tensor& x, int64_t mode, at::Tensor& hi, at::Tensor& lo) {
  TQB_ABI(a, x);
  TORCH_CHECK(x.dim() == 2 && x.scalar_type() == at::kComplexFloat && hi.scalar_type() == at::kFloat && lo.scalar_type() == at::kFloat,
              "tqb_gemm_prep: x complex64 [rows][cols], planes float32");
  const int64_t want = (mode == 1 ? 2 : 4) * x.numel();
  TORCH_CHECK(hi.numel() == want && lo.numel() == want, "tqb_gemm_prep: plane size");
  check(a, a.gemm_prep(reinterpret_cast<const float*>(x.data_ptr()), x.size(0), x.size(1), (int)mode, fpm(hi), fpm(lo), stream_of(x)),
        "tqb_gemm_prep");
}

void gemm_tf32x3(const at::Tensor& a_hi, const at::Tensor& a_lo, const at::Tensor& bt_hi, const at::Tensor& bt_lo, at::Tensor& c) {
  TQB_ABI(a, a_hi);
  TORCH_CHECK(a_hi.dim() == 2 && bt_hi.dim() == 2 && c.dim() == 2 && a_hi.size(1) == bt_hi.size(1) && c.size(0) == a_hi.size(0) &&
                  c.size(1) == bt_hi.size(0) && a_lo.sizes() == a_hi.sizes() && bt_lo.sizes() == bt_hi.sizes(),
              "tqb_gemm_tf32x3: C[M][N] = A[M][K] * Bt[N][K]^T");
  const int64_t M = a_hi.size(0), N = bt_hi.size(0), K = a_hi.size(1);
  const int64_t ws = a.gemm_workspace_floats(M, N, K);
  at::Tensor workspace;   // torch's caching allocator: device memory is plumbing
  if (ws > 0) workspace = at::empty({ws}, c.options());
  check(a, a.gemm_tf32x3(fp(a_hi), fp(a_lo), fp(bt_hi), fp(bt_lo), fpm(c), ws > 0 ? static_cast<float*>(workspace.data_ptr()) : nullptr,
                         M, N, K, stream_of(a_hi)), "tqb_gemm_tf32x3");
}

// HOST helper of the ABI (CPython's Mersenne Twister stream, csrc/tqb_host_rng.h): CPU tensors by design; served by
// whichever build of the ABI is bound (the stream is the host's in both)
void host_uniforms(at::Tensor& mt_state, int64_t n, at::Tensor& out) {
  const Abi& a = g_cuda.handle != nullptr ? g_cuda : g_emu;
  TORCH_CHECK(a.handle != nullptr, "torchquantum_b200: libtqb200.so is not loaded");
  TORCH_CHECK(!mt_state.is_cuda() && !out.is_cuda() && mt_state.is_contiguous() && out.is_contiguous() &&
                  mt_state.numel() == 625 && mt_state.element_size() == 4 && out.scalar_type() == at::kFloat && out.numel() >= n,
              "tqb_host_uniforms: expects a 625-word host state and a float32 host buffer of n elements");
  check(a, a.host_uniforms(static_cast<uint32_t*>(mt_state.data_ptr()), n, static_cast<float*>(out.data_ptr())),
        "tqb_host_uniforms");
}

}  // namespace

TORCH_LIBRARY(tqb200, m) {
  m.def("tqb_host_uniforms(Tensor(a!) mt_state, int n, Tensor(b!) out) -> ()");
  m.def("tqb_split_tf32(Tensor x, Tensor(a!) hi, Tensor(b!) lo) -> ()");
  m.def("tqb_gemm_prep(Tensor x, int mode, Tensor(a!) hi, Tensor(b!) lo) -> ()");
  m.def("tqb_gemm_tf32x3(Tensor a_hi, Tensor a_lo, Tensor bt_hi, Tensor bt_lo, Tensor(a!) c) -> ()");
  m.def("tqb_exchange_p2p(Tensor src, Tensor(a!) local_dst, int[] peer_ptrs, int n_local, int bsz, int[] victim_bits, "
        "int rank) -> ()");
  m.def("_bind(str path, bool emulated) -> int", &bind);
  m.def("tqb_build_mats(Tensor table, int n_records, int n_shared_records, int bsz, Tensor? params_sh, Tensor? params_ps, "
        "int ps_stride, Tensor? consts, Tensor(a!) msh, Tensor(b!) mps, int mps_stride) -> ()");
  m.def("tqb_build_mats_bwd(Tensor table, int n_records, int bsz, Tensor? params_sh, Tensor? params_ps, int ps_stride, "
        "Tensor gr, int gr_stride, Tensor(a!) gparams_sh_partial, int sh_stride, Tensor(b!) gparams_ps) -> ()");
  m.def("tqb_pass_forward(Tensor src, Tensor(a!) dst, int n_qubits, int bsz, Tensor prog, int[] desc, Tensor msh, Tensor mps, "
        "int mps_stride, int ext_hi) -> ()");
  m.def("tqb_pass_backward(Tensor psi_src, Tensor lam_src, Tensor(a!) psi_dst, Tensor(b!) lam_dst, int n_qubits, int bsz, "
        "Tensor prog, int[] desc, Tensor msh, Tensor mps, int mps_stride, Tensor(c!) gr, int gr_stride, int ext_hi) -> ()");
  m.def("tqb_apply_dense(Tensor src, Tensor(a!) dst, int n_qubits, int bsz, Tensor mats, int mat_stride_complex, int k, "
        "Tensor bitpos, int conj_transpose) -> ()");
  m.def("tqb_apply_dense_tc(Tensor src, Tensor(a!) dst, int n_qubits, int bsz, Tensor mat, int k, int[] bitpos_host, "
        "int conj_transpose) -> ()");
  m.def("tqb_block_build(Tensor blocks, int n_blocks, Tensor ops, Tensor consts, Tensor params_sh, int rank, "
        "Tensor(a!) a_planes) -> ()");
  m.def("tqb_apply_blocks(Tensor src, Tensor(a!) dst, Tensor(b!)? scratch, int n_qubits, int bsz, Tensor a_planes, Tensor blocks_host, "
        "int n_blocks, int renorm) -> ()");
  m.def("tqb_quadphase(Tensor(a!) state, int n_local, int bsz, Tensor coef, int n_bits, int rank) -> ()");
  m.def("tqb_dense_grad(Tensor psi_in, Tensor lam_out, int n_qubits, int bsz, int k, Tensor bitpos, Tensor(a!) G) -> ()");
  m.def("tqb_expval_pauli(Tensor state, int n_qubits, int bsz, Tensor xmask, Tensor zmask, Tensor ny, int n_terms, "
        "Tensor(a!) out) -> ()");
  m.def("tqb_expval_pauli_bwd(Tensor state, int n_qubits, int bsz, Tensor xmask, Tensor zmask, Tensor ny, int n_terms, "
        "Tensor grad_out, Tensor(a!) grad_state) -> ()");
  m.def("tqb_expval_z(Tensor state, int n_qubits, int bsz, Tensor(a!) out) -> ()");
  m.def("tqb_expval_z_bwd(Tensor state, int n_qubits, int bsz, Tensor grad_out, Tensor(a!) grad_state) -> ()");
  m.def("tqb_probs(Tensor state, int n_qubits, int bsz, Tensor(a!) probs) -> ()");
  m.def("tqb_sample_counts(Tensor probs, int n_qubits, int bsz, Tensor uniforms, int n_shots, Tensor(a!) cdf_scratch, "
        "Tensor(b!) counts) -> ()");
}

#define TQB_IMPL_ALL(m)                                  \
  m.impl("tqb_build_mats", &build_mats);                 \
  m.impl("tqb_build_mats_bwd", &build_mats_bwd);         \
  m.impl("tqb_pass_forward", &pass_forward);             \
  m.impl("tqb_pass_backward", &pass_backward);           \
  m.impl("tqb_apply_dense", &apply_dense);               \
  m.impl("tqb_apply_dense_tc", &apply_dense_tc);         \
  m.impl("tqb_block_build", &block_build);               \
  m.impl("tqb_apply_blocks", &apply_blocks);             \
  m.impl("tqb_quadphase", &quadphase);                   \
  m.impl("tqb_dense_grad", &dense_grad);                 \
  m.impl("tqb_expval_pauli", &expval_pauli);             \
  m.impl("tqb_expval_pauli_bwd", &expval_pauli_bwd);     \
  m.impl("tqb_expval_z", &expval_z);                     \
  m.impl("tqb_expval_z_bwd", &expval_z_bwd);             \
  m.impl("tqb_probs", &probs);                           \
  m.impl("tqb_sample_counts", &sample_counts);    \
  m.impl("tqb_exchange_p2p", &exchange_p2p);      \
  m.impl("tqb_split_tf32", &split_tf32);          \
  m.impl("tqb_gemm_prep", &gemm_prep);            \
  m.impl("tqb_gemm_tf32x3", &gemm_tf32x3);

TORCH_LIBRARY_IMPL(tqb200, CUDA, m) { TQB_IMPL_ALL(m) }
// CPU key: reaches only the injected host-emulation build (tests); abi_for() raises otherwise
TORCH_LIBRARY_IMPL(tqb200, CPU, m) {
  TQB_IMPL_ALL(m)
  m.impl("tqb_host_uniforms", &host_uniforms);
}
