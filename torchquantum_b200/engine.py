"""Execution engine: runs a planned circuit through the C ABI (libtqb200.so) and wires the
adjoint backward into torch autograd.

Forward of a fused stage:  parameters --tqb_build_mats--> per-sample matrix rows;
state --tqb_pass_forward (one launch per pass)--> new state.  Only the *output* state is kept for
the backward, which un-applies the gates in reverse (tqb_pass_backward) while accumulating
dL/dU, then maps dL/dU to dL/dtheta (tqb_build_mats_bwd): training memory is O(2^n) per sample
instead of the reference's O(depth * 2^n) autograd tape.
"""
import os

import numpy as np
import torch

from . import _codes as C  # noqa: F401  (re-exported for tools)
from . import _native
from . import gates as G
from . import plan as P

_PLAN_CACHE = {}
_DEV_CACHE = {}
STATS = {"launches": 0, "plans_built": 0}


def get_plan(n, key, grad_mode, n_total=None, leave_layout=False):
    k = (n, n_total, grad_mode, key) if not leave_layout else (n, n_total, grad_mode, key, "leave")
    plan = _PLAN_CACHE.get(k)
    if plan is None:
        plan = P.build_plan(n, key, grad_mode, n_total, leave_layout=leave_layout)
        _PLAN_CACHE[k] = plan
        STATS["plans_built"] += 1
        if len(_PLAN_CACHE) > 512:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    return plan


def _dev(obj, attr, device):
    """numpy array attribute of a plan object -> cached device tensor."""
    k = (id(obj), attr, str(device))
    t = _DEV_CACHE.get(k)
    if t is None:
        t = torch.from_numpy(np.ascontiguousarray(getattr(obj, attr))).to(device)
        _DEV_CACHE[k] = (t, obj)  # keep obj alive so id() stays unique
        if len(_DEV_CACHE) > 8192:
            _DEV_CACHE.pop(next(iter(_DEV_CACHE)))
        return t
    return t[0]


def _desc(pp):
    if pp._desc_c is None:
        pp._desc_c = [int(v) for v in pp.desc]
    return pp._desc_c


def _build_mats(stage, bsz, psh, pps, externs, device):
    msh = torch.empty(stage.M_sh, dtype=torch.float32, device=device)
    mps = torch.empty(bsz, stage.M_ps, dtype=torch.float32, device=device)
    if stage.table.shape[0]:
        table = _dev(stage, "table", device)
        consts = _dev(stage, "consts", device)
        _native.call("tqb_build_mats", table, table.shape[0], stage.n_sh_rec, bsz, psh, pps,
                     pps.shape[1] if pps is not None else 0, consts, msh, mps, stage.M_ps)
        STATS["launches"] += 1
    for (rec, moff, dim, inverse, per_sample, goff), u in zip(stage.externs, externs):
        m = u.detach().to(torch.complex64)
        if inverse:
            m = m.conj().transpose(-1, -2)
        m = torch.view_as_real(m.contiguous()).reshape(-1, 2 * dim * dim)
        if per_sample:
            mps[:, moff:moff + 2 * dim * dim] = m
        else:
            msh[moff:moff + 2 * dim * dim] = m[0]
    return msh, mps


def _ext_hi(rank, n, pp):
    """Sharded statevector: the rank index supplies the index bits above the local 2^n amplitudes; the
    kernel sees them as the top bits of the compressed non-tile index (csrc/tqb_program.h, CM_O / PM_O)."""
    return (rank << (n - pp.T)) & 0xFFFFFFFF


def run_fused_inplace(state, stage, n, rank, psh, pps, externs):
    """No-grad execution of a fused stage IN PLACE on `state` (every pass reads and writes the same
    buffer: a tile is fully staged on chip before it is written back).  Used by the sharded device,
    where a second copy of a 16-64 GiB shard is not affordable."""
    _native.check_device(state)
    assert state.is_contiguous() and state.data_ptr() % 16 == 0
    bsz, dev = state.shape[0], state.device
    msh, mps = _build_mats(stage, bsz, psh, pps, externs, dev)
    for pp in stage.passes:
        prog = _dev(pp, "prog", dev)
        _native.call("tqb_pass_forward", state, state, n, bsz, prog, _desc(pp), msh,
                     mps, stage.M_ps, _ext_hi(rank, n, pp))
        STATS["launches"] += 1
    return state


def run_blocks(state, stage, n, rank, psh, inplace, timing=None, scratch=None):
    """Forward-only block stage (plan.BlockStage): the block unitaries are built on the device from the live parameters,
    then every block is one tensor-core pass.  Blocks that rotate the layout (TQB_KBF_ROTATE) are out of place: the state
    alternates between two buffers (`state` itself when `inplace`, and `scratch` or a fresh one), so the result may come
    back in either; the others work where the state is.
    timing: a list -> the blocks are launched one ABI call each and (start, end) CUDA events of every launch are
    appended (bench.py's per-launch roofline)."""
    _native.check_device(state)
    state = state.contiguous()
    if state.data_ptr() & 15:
        state, inplace = state.clone(), True
    bsz, dev = state.shape[0], state.device
    blocks_dev, ops_dev, consts = _dev(stage, "blocks", dev), _dev(stage, "ops", dev), _dev(stage, "consts", dev)
    planes = torch.empty(stage.n_blocks, 2, 128, 64, dtype=torch.float32, device=dev)
    _native.call("tqb_block_build", blocks_dev, stage.n_blocks, ops_dev, consts, psh, rank,
                 planes)
    rot = [bool(int(f) & C.KBF_ROTATE) for f in stage.blocks[:, C.KB_FLAGS]]
    n_rot = sum(rot)
    # two buffers a / b; the first block reads `state`; the result is in `out`
    a = state if inplace else torch.empty_like(state)
    need_b = n_rot > 0 and (inplace or n_rot - rot[0] > 0 or len(stage.items) > 1)
    if scratch is not None and (scratch.data_ptr() == state.data_ptr() or scratch.shape != state.shape
                                or not scratch.is_contiguous()):
        scratch = None
    b = (scratch if scratch is not None else torch.empty_like(state)) if need_b else None
    cur = state
    psh64 = None
    for item in stage.items:
        if item[0] == "quad":
            # quadratic-phase pass: coefficients (turns) = base + W . params, assembled on the device in float64
            q = stage.quads[item[1]]
            coef = _dev(q, "base", dev)
            if len(q.idx):
                if psh64 is None:
                    psh64 = psh.to(torch.float64)
                coef = coef.index_add(0, _dev(q, "idx", dev), _dev(q, "w", dev) * psh64[_dev(q, "pidx", dev)])
            words = ((coef - torch.floor(coef)) * 4294967296.0).to(torch.int64).bitwise_and_(0xFFFFFFFF).to(torch.int32)
            if cur is state and not inplace:     # the caller's state must survive: the pass works on a copy
                a.copy_(state)
                cur = a
            _native.call("tqb_quadphase", cur, n, bsz, words, q.n_bits, rank)
            STATS["launches"] += 1
            continue
        _, first, count = item
        if timing is not None:
            for k in range(first, first + count):
                if rot[k]:
                    dst = b if cur is a else a
                else:
                    dst = cur if (cur is not state or inplace) else a
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _native.call("tqb_apply_blocks", cur, dst, None, n, bsz, planes[k],
                             torch.from_numpy(stage.blocks[k:k + 1]), 1, 1)
                e1.record()
                timing.append((e0, e1))
                cur = dst
            continue
        # one ABI call for the run of blocks: rotating blocks alternate between the two buffers, the rest work where the
        # state is; the call leaves the result in `dst` (an odd number of moves ends in the other buffer)
        r = sum(rot[first:first + count])
        if cur is state and not inplace:
            dst, other = (a, b)                          # first write goes to a fresh buffer, `state` is only read
        elif r % 2:
            dst, other = (b if cur is a else a), cur     # the buffer the state leaves serves as the scratch
        else:
            dst, other = cur, (b if cur is a else a)
        _native.call("tqb_apply_blocks", cur, dst, other, n, bsz, planes[first:first + count],
                     torch.from_numpy(stage.blocks[first:first + count]), count, 1)
        cur = dst
    STATS["launches"] += 1 + stage.n_blocks
    return cur


class FusedFn(torch.autograd.Function):
    """state [bsz, 2^n] complex64 -> state, through all passes of one fused stage."""

    @staticmethod
    def forward(ctx, state, psh, pps, stage, n, rank, *externs):
        _native.check_device(state)
        state = state.contiguous()
        if state.data_ptr() & 15:   # tiles are filled with 16-byte async copies
            state = state.clone()
        bsz = state.shape[0]
        dev = state.device
        msh, mps = _build_mats(stage, bsz, psh, pps, externs, dev)
        out = torch.empty_like(state)
        src = state
        for pp in stage.passes:
            prog = _dev(pp, "prog", dev)
            _native.call("tqb_pass_forward", src, out, n, bsz, prog, _desc(pp), msh,
                         mps, stage.M_ps, _ext_hi(rank, n, pp))
            STATS["launches"] += 1
            src = out
        if not stage.passes:
            out.copy_(state)
        ctx.stage, ctx.n, ctx.rank = stage, n, rank
        ctx.n_ext = len(externs)
        ctx.save_for_backward(state if stage.save_input else out, psh, pps, msh, mps, *externs)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        stage, n, rank = ctx.stage, ctx.n, ctx.rank
        saved = ctx.saved_tensors
        psi0, psh, pps, msh, mps = saved[:5]
        externs = saved[5:]
        bsz = psi0.shape[0]
        dev = psi0.device
        grad_out = grad_out.contiguous()
        multi = any(pp.T < n for pp in stage.passes)
        gr = (torch.zeros if multi else torch.empty)(bsz, stage.G, dtype=torch.float32, device=dev)
        psi = torch.empty_like(psi0)
        lam = torch.empty_like(psi0)
        psrc, lsrc = psi0, grad_out
        for pp in reversed(stage.passes):
            prog = _dev(pp, "prog", dev)
            _native.call("tqb_pass_backward", psrc, lsrc, psi, lam, n, bsz, prog,
                         _desc(pp), msh, mps, stage.M_ps, gr, stage.G, _ext_hi(rank, n, pp))
            STATS["launches"] += 1
            psrc, lsrc = psi, lam
        if not stage.passes:
            lam.copy_(grad_out)
        gsh = gps = None
        if stage.any_grad and (stage.P_sh or stage.P_ps):
            table = _dev(stage, "table", dev)
            gsh_part = torch.zeros(bsz, max(stage.P_sh, 1), dtype=torch.float32, device=dev)
            gps_full = torch.zeros(bsz, max(stage.P_ps, 1), dtype=torch.float32, device=dev)
            _native.call("tqb_build_mats_bwd", table, table.shape[0], bsz, psh, pps,
                         pps.shape[1], gr, stage.G, gsh_part, gsh_part.shape[1],
                         gps_full)
            STATS["launches"] += 1
            if stage.P_sh and ctx.needs_input_grad[1]:
                gsh = gsh_part.sum(0)
            if stage.P_ps and ctx.needs_input_grad[2]:
                gps = gps_full
        gext = []
        for i, ((rec, moff, dim, inverse, per_sample, goff), u) in enumerate(zip(stage.externs, externs)):
            if not ctx.needs_input_grad[6 + i]:
                gext.append(None)
                continue
            g = torch.view_as_complex(gr[:, goff:goff + 2 * dim * dim].reshape(bsz, dim, dim, 2).contiguous())
            if inverse:  # applied A = U^H  =>  dL/dU = (dL/dA)^H
                g = g.conj().transpose(-1, -2)
            if not per_sample:
                g = g.sum(0, keepdim=True)
            gext.append(g.reshape(u.shape))
        gstate = lam if ctx.needs_input_grad[0] else None
        return (gstate, gsh, gps, None, None, None, *gext)


DENSE_TC = True   # tensor-core (3xTF32) dense kernel for one shared matrix on 4-7 wires of a state of >= 12 qubits


def _dense_apply(src, dst, n, bsz, m, k, bitpos, bp_dev, conj_t):
    """dst = U src (or U^H src) for the k-qubit matrix m on bit positions `bitpos` (matrix MSB first)."""
    if DENSE_TC and m.dim() == 2 and 4 <= k <= 7 and n >= 12 and not _native.is_emulated():
        host = [int(b) for b in bitpos]
        _native.call("tqb_apply_dense_tc", src, dst, n, bsz, m, k, host, 1 if conj_t else 0)
    else:
        _native.call("tqb_apply_dense", src, dst, n, bsz, m, (4 ** k) if m.dim() == 3 else 0,
                     k, bp_dev, 1 if conj_t else 0)
    STATS["launches"] += 1


class DenseFn(torch.autograd.Function):
    """Generic k-qubit dense gate (qft, qubitunitary on >= 3 wires), out of place."""

    @staticmethod
    def forward(ctx, state, mat, bitpos, n, inverse):
        _native.check_device(state)
        state = state.contiguous()
        bsz = state.shape[0]
        m = mat.to(torch.complex64).contiguous()
        k = len(bitpos)
        out = torch.empty_like(state)
        bp = torch.tensor(bitpos, dtype=torch.int32, device=state.device)
        _dense_apply(state, out, n, bsz, m, k, bitpos, bp, inverse)
        ctx.save_for_backward(state, m, bp)
        ctx.n, ctx.inverse, ctx.k, ctx.bitpos = n, inverse, k, tuple(bitpos)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        state, m, bp = ctx.saved_tensors
        n, inverse, k = ctx.n, ctx.inverse, ctx.k
        bsz = state.shape[0]
        grad_out = grad_out.contiguous()
        gstate = gmat = None
        if ctx.needs_input_grad[0]:
            gstate = torch.empty_like(state)
            _dense_apply(grad_out, gstate, n, bsz, m, k, ctx.bitpos, bp, not inverse)
        if ctx.needs_input_grad[1]:
            g = torch.empty(bsz, 2 ** k, 2 ** k, dtype=torch.complex64, device=state.device)
            _native.call("tqb_dense_grad", state, grad_out, n, bsz, k, bp, g)
            STATS["launches"] += 1
            if inverse:
                g = g.conj().transpose(-1, -2)
            gmat = g.sum(0) if m.dim() == 2 else g
        return gstate, gmat, None, None, None


def _reset(state, n, wires):
    """functional/reset.py:19-42 — zero the |1> slice of each wire, renormalise
    (util/utils.py:789-814).  Non-unitary and rare: plain torch ops on the device tensor."""
    s = state.reshape([state.shape[0]] + [2] * n).clone()
    for w in wires:
        idx = [slice(None)] * s.dim()
        idx[w + 1] = 1
        s[tuple(idx)] = 0
    flat = s.reshape(state.shape[0], -1)
    energy = (flat.abs() ** 2).sum(-1)
    fix = torch.zeros_like(flat)
    fix[:, 0] = (energy == 0).to(flat.dtype)
    flat = flat + fix
    return flat * torch.sqrt(1 / (flat.abs() ** 2).sum(-1)).unsqueeze(-1)


class RecordChunk(list):
    """A run of gate records that a layer template queues unchanged step after step (same Parameter
    objects, same wires).  The per-record plan keys are cached on the chunk, so a step only re-checks
    the `requires_grad` flags instead of rebuilding several hundred key tuples."""

    def __init__(self, records):
        super().__init__(records)
        self.params = [r[2] for r in records if r[2] is not None]
        self._keys = {}

    def collapsible(self):
        """every record is the same for all samples (constants / shared [1, p] parameters, unitary): see run_circuit"""
        c = self.__dict__.get("_collapsible")
        if c is None:
            c = self.__dict__["_collapsible"] = all(_collapsible(r) for r in self)
        return c

    def key_for(self, bsz, grad_mode):
        flags = [p.requires_grad for p in self.params]
        hit = self._keys.get((bsz, grad_mode))
        if hit is not None and hit[0] == flags:
            return hit[1]
        key = tuple([_record_key(r, bsz, grad_mode) for r in self])
        self._keys[(bsz, grad_mode)] = (flags, key)
        return key


def _record_key(record, bsz, grad_mode):
    name, wires, params, inverse = record
    pmode, grad, ext_dim = None, False, 0
    if params is not None:
        grad = grad_mode and params.requires_grad
        if name in G.UNITARY_LIKE:
            pmode, ext_dim = ("extps" if params.dim() == 3 and params.shape[0] == bsz and bsz > 1 else "ext"), \
                params.shape[-1]
        else:
            if params.shape[0] == 1:
                pmode = "sh"
            elif params.shape[0] == bsz:
                pmode = "ps"
            else:
                raise RuntimeError(
                    f"{name}: parameter batch {params.shape[0]} does not match state batch {bsz}")
            if G.SPECS[name].n_params != params.shape[1]:
                raise RuntimeError(f"{name}: expected {G.SPECS[name].n_params} parameters, got {params.shape[1]}")
    return (name, wires, inverse, pmode, grad, ext_dim)


# ---------------------------------------------------------------------------------------
# Batch collapse: a long run of gates that are the SAME for every sample (constants and shared [1, p] parameters) acting
# on a batch larger than the Hilbert space is ONE 2^n x 2^n unitary W: it is built once by running the run on the
# identity batch (2^n "samples" instead of bsz), and applied to the batch as a complex GEMM on the tensor cores.  The
# reference's static mode does the same thing block-wise on the host (graph/graphs.py:713-861: unitary of the block,
# then one bmm for the whole batch).  Autograd: dL/dW comes out of the GEMM's backward (psi^H . lambda, a GEMM again) and
# flows into the adjoint sweep of the run over 2^n columns instead of bsz samples, so the backward shrinks by the same
# factor bsz / 2^n.  Config 2 (n = 10, bsz = 4096, 600 shared gates): 4x fewer gate-amplitude updates in both directions.
# ---------------------------------------------------------------------------------------
COLLAPSE = os.environ.get("TQB200_COLLAPSE", "1") != "0"
COLLAPSE_MIN_GATES = 64      # shorter runs are cheaper through the register-tile passes than a W build + GEMM
COLLAPSE_MIN_N = 7           # the GEMM tiles are 128 x 256 x 32 real: 2 * 2^n >= 256
COLLAPSE_MAX_N = 12          # W is 2^n x 2^n complex64: 128 MiB at n = 12
COLLAPSE_MIN_BATCH_RATIO = 2  # bsz >= ratio * 2^n
_EYE_CACHE = {}


def _identity_batch(n, device):
    k = (n, str(device))
    t = _EYE_CACHE.get(k)
    if t is None:
        t = _EYE_CACHE[k] = torch.eye(2 ** n, dtype=torch.complex64, device=device)
    return t


def _collapsible(record):
    name, wires, params, inverse = record
    if name in G.UNITARY_LIKE or G.SPECS[name].kind == G.RESET or not G.SPECS[name].unitary:
        return False
    return params is None or (params.dim() == 2 and params.shape[0] == 1)


class ComplexGemmFn(torch.autograd.Function):
    """out[b, :] = sum_c a[b, c] w[c, :]   (a: [bsz, K], w: [K, N], complex64) with the conjugate-transposed GEMMs as its
    backward."""

    @staticmethod
    def forward(ctx, a, w):
        ctx.save_for_backward(a, w)
        return complex_gemm(a, w)

    @staticmethod
    def backward(ctx, g):
        a, w = ctx.saved_tensors
        g = g.contiguous()
        ga = complex_gemm(g, w.conj().transpose(0, 1).contiguous()) if ctx.needs_input_grad[0] else None
        gw = complex_gemm(a, g, a_h=True) if ctx.needs_input_grad[1] else None
        return ga, gw


def _tf32_planes(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _native.call("tqb_split_tf32", x, hi, lo)
    return hi, lo


def _prep_planes(x, mode):
    """tqb_gemm_prep: the GEMM operand of x^H (mode 1: A) or of x (mode 2: Bt) as TF32 hi / lo planes, one kernel"""
    rows, cols = x.shape
    shape = (cols, 2 * rows) if mode == 1 else (2 * cols, 2 * rows)
    hi = torch.empty(shape, dtype=torch.float32, device=x.device)
    lo = torch.empty(shape, dtype=torch.float32, device=x.device)
    _native.call("tqb_gemm_prep", x, mode, hi, lo)
    return hi, lo


def complex_gemm(a, b, a_h=False):
    """a [M, K] @ b [K, N] (a_h: a^H @ b with a [K, M]), complex64, on the tensor cores (tqb_gemm_tf32x3: 3xTF32, fp32
    accuracy).  The complex product is a real GEMM on the interleaved (re, im) views: A = a as [M][2K], Bt[2n + w][2k + v]
    = the real 2x2 embedding of b[k][n].  Tile-aligned shapes take their operands from ONE fused preparation kernel each
    (tqb_gemm_prep: conjugate / transpose / embed / TF32 split in a single pass)."""
    _native.check_device(a)
    if a_h:
        K, M = a.shape
    else:
        M, K = a.shape
    N = b.shape[1]
    if K % 32 == 0 and M % 128 == 0 and N % 128 == 0:
        a, b = a.contiguous(), b.contiguous()
        if a_h:
            a_hi, a_lo = _prep_planes(a, 1)
        else:
            a_hi, a_lo = _tf32_planes(torch.view_as_real(a).reshape(M, 2 * K))
        b_hi, b_lo = _prep_planes(b, 2)
        c = torch.empty(M, 2 * N, dtype=torch.float32, device=a.device)
        _native.call("tqb_gemm_tf32x3", a_hi, a_lo, b_hi, b_lo, c)
        STATS["launches"] += 3
        return torch.view_as_complex(c.view(M, N, 2))
    if a_h:
        a = a.conj().transpose(0, 1).contiguous()
    Kp = -(-K // 16) * 16            # tile sizes of the kernel: 128 rows, 32 real K, 256 real N (zero padding is exact)
    if Kp != K:
        a = torch.nn.functional.pad(a, (0, Kp - K))
        b = torch.nn.functional.pad(b, (0, 0, 0, Kp - K))
        K = Kp
    assert (2 * N) % 256 == 0, "complex_gemm: N must be a multiple of 128"
    ar = torch.view_as_real(a.contiguous()).reshape(M, 2 * K)
    Mp = -(-M // 128) * 128
    if Mp != M:
        ar = torch.nn.functional.pad(ar, (0, 0, 0, Mp - M))
    bt = b.transpose(0, 1).conj()                                   # [N, K]: rows 2n of Bt are (re, -im) of b[:, n]
    btr = torch.view_as_real(torch.stack([bt, 1j * bt], dim=1).contiguous()).reshape(2 * N, 2 * K)   # rows 2n + 1: (im, re)
    a_hi, a_lo = _tf32_planes(ar)
    b_hi, b_lo = _tf32_planes(btr)
    c = torch.empty(Mp, 2 * N, dtype=torch.float32, device=a.device)
    _native.call("tqb_gemm_tf32x3", a_hi, a_lo, b_hi, b_lo, c)
    STATS["launches"] += 3
    return torch.view_as_complex(c[:M].view(M, N, 2))


def _run_collapsed(state, n, records):
    """records: a run of batch-independent gates -> state @ W^T with W^T[c, :] = circuit |c>"""
    wt = run_circuit(_identity_batch(n, state.device), n, records, _no_collapse=True)
    return ComplexGemmFn.apply(state.contiguous(), wt)


def run_circuit(state, n, queue, n_total=None, rank=0, inplace=False, _no_collapse=False, scratch=None, leave_layout=False):
    """state: [bsz, 2^n] complex64; queue: list of records (name, wires, params, inverse) with canonical
    names and normalised params (2-D float tensors, or complex matrices for unitary-like gates), or
    RecordChunks of such records.
    n_total / rank: sharded statevector (plan.build_plan); inplace: overwrite `state` (no-grad only).
    scratch: a second state-sized buffer the caller owns (with inplace): block stages that rotate the qubit layout
    alternate between the two, so the result may come back in either.
    leave_layout: the caller tracks the qubit layout itself (sharded device): -> (state, final_pos) where final_pos is
    None or the logical -> physical bit map the last (rotating block) stage left behind instead of sorting the qubits home."""
    bsz = state.shape[0]
    chunked = any(type(item) is RecordChunk for item in queue)
    if chunked:
        records = []
        for item in queue:
            if type(item) is RecordChunk:
                records.extend(item)
            else:
                records.append(item)
    else:
        records = queue
    if (COLLAPSE and not _no_collapse and n_total is None and COLLAPSE_MIN_N <= n <= COLLAPSE_MAX_N
            and len(records) >= COLLAPSE_MIN_GATES
            and bsz >= COLLAPSE_MIN_BATCH_RATIO * 2 ** n):
        # maximal runs of batch-independent queue items (single records, or whole RecordChunks of a layer template, whose
        # cached plan keys keep working); long runs go through W, everything else runs as before
        runs = []    # [flag, n_gates, items]
        for item in queue:
            if type(item) is RecordChunk:
                flag, count = item.collapsible(), len(item)
            else:
                flag, count = _collapsible(item), 1
            if runs and runs[-1][0] == flag:
                runs[-1][1] += count
                runs[-1][2].append(item)
            else:
                runs.append([flag, count, [item]])
        out, pending = state, []
        for flag, count, items in runs:
            if flag and count >= COLLAPSE_MIN_GATES:
                if pending:
                    out, pending = run_circuit(out, n, pending, _no_collapse=True), []
                out = _run_collapsed(out, n, items)
            else:
                pending += items   # short batch-independent runs stay with their per-sample neighbours
        if pending:
            out = run_circuit(out, n, pending, _no_collapse=True)
        return out
    grad_mode = torch.is_grad_enabled() and (
        state.requires_grad or any(r[2] is not None and r[2].requires_grad for r in records))
    if chunked:
        key = ()
        for item in queue:
            key += item.key_for(bsz, grad_mode) if type(item) is RecordChunk else (_record_key(item, bsz, grad_mode),)
    else:
        key = [_record_key(r, bsz, grad_mode) for r in records]
    plan = get_plan(n, tuple(key), grad_mode, n_total, leave_layout and not grad_mode)
    for stage in plan.stages:
        if stage.kind == "fused":
            dev = state.device
            if stage.sh_recs:
                # (2-D [1, p] parameters are concatenated as they are: a reshape per gate would add one autograd
                # node per parameter to every step)
                psh = torch.cat([records[r][2] if records[r][2].dim() == 2 else records[r][2].reshape(1, -1)
                                 for r in stage.sh_recs], dim=1).reshape(-1)
                psh = psh.to(device=dev, dtype=torch.float32)
            else:
                psh = torch.zeros(1, dtype=torch.float32, device=dev)
            if stage.ps_recs:
                pps = torch.cat([records[r][2] for r in stage.ps_recs], dim=1).to(device=dev, dtype=torch.float32)
                pps = pps.contiguous()
            else:
                pps = torch.zeros(bsz, 1, dtype=torch.float32, device=dev)
            externs = [records[e[0]][2].to(dev) for e in stage.externs]
            if inplace and not grad_mode:
                state = run_fused_inplace(state, stage, n, rank, psh, pps, externs)
            else:
                state = FusedFn.apply(state, psh, pps, stage, n, rank, *externs)
        elif stage.kind == "blocks":
            dev = state.device
            if stage.sh_recs:
                psh = torch.cat([records[r][2].reshape(1, -1) for r in stage.sh_recs], dim=1).reshape(-1)
                psh = psh.detach().to(device=dev, dtype=torch.float32)
            else:
                psh = torch.zeros(1, dtype=torch.float32, device=dev)
            prev = state
            state = run_blocks(state, stage, n, rank, psh, inplace, scratch=scratch)
            if inplace and state.data_ptr() != prev.data_ptr():
                scratch = prev     # the state moved to the other buffer: the one it left is the spare now
        elif stage.kind == "dense":
            name, wires, params, inverse = records[stage.rec]
            if stage.const is not None:
                mat = torch.from_numpy(stage.const).to(state.device)
            else:
                mat = params.to(state.device)
            state = DenseFn.apply(state, mat, tuple(stage.bitpos), n, inverse)
        else:
            state = _reset(state, n, stage.wires)
    return (state, plan.final_pos) if leave_layout else state
