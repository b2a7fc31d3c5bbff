"""Host planner: turns a queued gate list into fused passes for the sm_100a kernels.

This replaces the reference's static-mode scheduler (graph/graphs.py:384-509, greedy blocks of
`wires_per_block` qubits multiplied into dense unitaries) with a planner shaped by the memory
system instead: a *pass* is one HBM round trip of the state with a tile of 2^T amplitudes staged
in shared memory; inside a pass, *groups* of gates whose targets fit R register bits are applied
without touching shared memory.  Gates on disjoint qubits commute, so both levels use the same
dependency-respecting greedy packing.  The output is the int32 program of csrc/tqb_program.h.
"""
import functools

import numpy as np

from . import _codes as C
from . import gates as G

T_SINGLE = 13   # largest state handled as one tile per sample (2^13 amplitudes = 64 KiB)
T_MULTI = 12    # tile bits when the state is streamed through HBM
L_MIN = 5       # contiguous low bits always in the tile (2^5 amplitudes = 256 B runs)
R_MAX = C.MAX_R


class LOp:
    """A lowered micro-op on global bit positions."""
    __slots__ = ("type", "targets", "pmask", "cmask", "cval", "nfloat", "kind", "inverse", "pmode",
                 "poff", "nparam", "const", "rec", "grad", "nounapply", "qmask", "ext_dim", "_lt")

    def __init__(self):
        self.pmask = self.cmask = self.cval = 0
        self.targets = ()
        self.kind, self.inverse, self.pmode, self.poff, self.nparam = C.K_CONST, False, None, 0, 0
        self.const, self.rec, self.grad, self.nounapply, self.ext_dim = None, -1, False, False, 0


def _bits(n, wires):
    return [n - 1 - int(w) for w in wires]


def _const_floats(mat, inverse):
    m = np.asarray(mat, dtype=np.complex64)
    if inverse:
        m = m.conj().T if m.ndim == 2 else m.conj()
    return np.ascontiguousarray(m).view(np.float32).reshape(-1)


def lower(n, name, wires, inverse, pmode, grad, rec, ext_dim=0):
    """One queue record -> LOp (fusable) or None for DENSE / RESET / NOP kinds.
    pmode: None | 'sh' | 'ps' | 'ext' (explicit matrix supplied by the caller)."""
    spec = G.SPECS[name]
    op = LOp()
    op.rec, op.grad, op.inverse, op.pmode = rec, grad, inverse, pmode
    bits = _bits(n, wires)
    kind = spec.kind
    const = spec.const
    if name == "qft" and len(wires) <= 2:
        kind = G.U1Q if len(wires) == 1 else G.U2Q
        const = G.qft_matrix_np(len(wires))
    if pmode == "ext":
        if ext_dim == 2:
            kind = G.U1Q
        elif ext_dim == 4:
            kind = G.U2Q
        else:
            return None
        op.kind, op.ext_dim = C.K_EXTERN, ext_dim
    elif kind in (G.DENSE, G.RESET, G.NOP):
        return None
    n_ctrl = spec.n_ctrl if spec.n_ctrl >= 0 else len(wires) - 1
    if pmode == "ext":
        n_ctrl = 0
    for b in bits[:n_ctrl]:
        op.cmask |= 1 << b
    op.cval = 0 if spec.neg_ctrl else op.cmask
    rest = bits[n_ctrl:]
    if kind == G.U1Q:
        assert len(rest) == 1, f"{name}: expected one target wire"
        op.type, op.targets, op.nfloat = C.T_U1Q, (rest[0],), 8
    elif kind == G.U2Q:
        assert len(rest) == 2, f"{name}: expected two target wires"
        op.type, op.targets, op.nfloat = C.T_U2Q, (rest[0], rest[1]), 32
    else:
        op.type, op.nfloat = C.T_DIAG, 4
        for b in rest:
            op.pmask |= 1 << b
    if op.kind != C.K_EXTERN:
        if spec.builder is not None and pmode in ("sh", "ps"):
            op.kind, op.nparam = spec.builder, spec.n_params
        else:
            op.kind, op.const = C.K_CONST, _const_floats(const, inverse)
    tm = 0
    for t in op.targets:
        tm |= 1 << t
    op.qmask = tm | op.pmask | op.cmask
    assert bin(op.qmask).count("1") == len(set(bits)), f"{name}: repeated wires {wires}"
    return op


# ---------------------------------------------------------------------------------------
# greedy, dependency-respecting packing
# ---------------------------------------------------------------------------------------
def _pack_passes(lops, L, cap):
    """Split into passes: each pass may hold ops whose targets >= L fit in `cap` free tile bits."""
    remaining, passes = list(lops), []
    while remaining:
        chosen, blocked, part, rest = set(), 0, [], []
        for op in remaining:
            need = {t for t in op.targets if t >= L and t not in chosen}
            if not (op.qmask & blocked) and len(chosen) + len(need) <= cap:
                part.append(op)
                chosen |= need
            else:
                blocked |= op.qmask
                rest.append(op)
        passes.append((chosen, part))
        remaining = rest
    return passes


def _pack_groups(ops, R):
    """Split a pass's ops (tile-local targets in op._lt) into register groups of <= R target bits.
    U2Q ops need their targets at register indices (1, 0)."""
    remaining, groups = list(ops), []
    while remaining:
        reg, blocked, part, rest = [], 0, [], []
        for op in remaining:
            ok = not (op.qmask & blocked)
            if ok:
                if op.type == C.T_U1Q:
                    t = op._lt[0]
                    if t not in reg:
                        if len(reg) < R:
                            reg.append(t)
                        else:
                            ok = False
                elif op.type == C.T_U2Q:
                    tm, tl = op._lt
                    if not reg and R >= 2:
                        reg.extend([tl, tm])
                    elif not (len(reg) >= 2 and reg[0] == tl and reg[1] == tm):
                        ok = False
            if ok:
                part.append(op)
            else:
                blocked |= op.qmask
                rest.append(op)
        groups.append((reg, part))
        remaining = rest
    return groups


def _swz(idx):
    return idx ^ ((idx >> 4) & 15)


@functools.lru_cache(maxsize=4096)
def _fill_registers(T, R, reg):
    """Complete `reg` (tuple of tile-local target bits) to R distinct bits, choosing the filler bits
    that minimise shared-memory bank conflicts of the group's 64-bit accesses."""
    reg = list(reg)
    free = [b for b in range(T) if b not in reg]
    missing = R - len(reg)
    if missing == 0 and T - R < 4:
        return tuple(reg)
    import itertools

    def degree(regset):
        if T - R < 4:
            return 1
        sp = sorted(regset)
        worst = 1
        for half in (0, 16):
            banks = {}
            for lane in range(16):
                w = lane + half
                if w >= (1 << (T - R)):
                    break
                for p in sp:
                    w = ((w >> p) << (p + 1)) | (w & ((1 << p) - 1))
                k = _swz(w) & 15
                banks[k] = banks.get(k, 0) + 1
            worst = max(worst, max(banks.values()))
        return worst

    if missing == 0:
        return tuple(reg)
    best, best_deg = None, None
    for combo in itertools.combinations(free, missing):
        d = degree(reg + list(combo))
        if best is None or d < best_deg:
            best, best_deg = combo, d
            if d == 1:
                break
    return tuple(reg + list(best))


class PassProgram:
    __slots__ = ("T", "R", "L", "prog", "n_ops", "n_groups", "n_seg", "tile_bits")


class FusedStage:
    """One autograd node: a list of passes sharing one matrix row layout."""
    kind = "fused"

    def __init__(self):
        self.passes = []
        self.table = None      # int32 [nrec, 8]
        self.consts = None     # float32
        self.M = 0             # floats per matrix row
        self.P_sh = self.P_ps = 0
        self.sh_recs, self.ps_recs = [], []   # queue indices in parameter order
        self.externs = []      # (rec, moff, dim, inverse)
        self.save_input = False
        self.any_grad = False
        self.n_ops = 0


class DenseStage:
    kind = "dense"

    def __init__(self, rec, bitpos, const, inverse, grad):
        self.rec, self.bitpos, self.const, self.inverse, self.grad = rec, bitpos, const, inverse, grad


class ResetStage:
    kind = "reset"

    def __init__(self, wires):
        self.wires = wires


def _i32(v):
    return int(np.array(v & 0xFFFFFFFF, dtype=np.uint32).view(np.int32))


def build_fused_stage(lops, n_local, n_total=None, t_single=None, t_multi=None, l_min=None):
    """lops: LOps in program order (targets < n_local; masks may use bits up to n_total)."""
    n_total = n_local if n_total is None else n_total
    t_single = T_SINGLE if t_single is None else t_single
    t_multi = T_MULTI if t_multi is None else t_multi
    l_min = L_MIN if l_min is None else l_min
    st = FusedStage()
    st.n_ops = len(lops)
    # parameter slots in queue order
    for op in lops:
        if op.pmode == "sh":
            op.poff = st.P_sh
            st.P_sh += op.nparam
            st.sh_recs.append(op.rec)
        elif op.pmode == "ps":
            op.poff = st.P_ps
            st.P_ps += op.nparam
            st.ps_recs.append(op.rec)
        st.any_grad |= op.grad
    # passes
    t_single = max(t_single, min(t_multi, n_local))
    if n_local <= t_single:
        raw = [(set(), list(lops))]
        T, L = n_local, n_local
    else:
        T, L = t_multi, l_min
        raw = _pack_passes(lops, L, T - L)
    R = min(R_MAX, T)
    consts, table, mat_off = [], [], 0
    const_off = 0
    for chosen, ops in raw:
        if n_local <= t_single:
            tile_bits = list(range(n_local))
        else:
            tile = set(range(L)) | chosen
            b = L
            while len(tile) < T:   # pad with the lowest unused bits: longer contiguous runs
                if b not in tile:
                    tile.add(b)
                b += 1
            tile_bits = sorted(tile)
        Lp = 0
        while Lp < T and tile_bits[Lp] == Lp:
            Lp += 1
        loc = {g: i for i, g in enumerate(tile_bits)}
        outer = [p for p in range(n_total) if p not in loc]
        oloc = {g: i for i, g in enumerate(outer)}

        def split(mask):
            t = o = 0
            p = 0
            while mask >> p:
                if (mask >> p) & 1:
                    if p in loc:
                        t |= 1 << loc[p]
                    else:
                        o |= 1 << oloc[p]
                p += 1
            return t, o

        for op in ops:
            op._lt = tuple(loc[t] for t in op.targets)
        groups = _pack_groups(ops, R)
        # split oversize groups, then cut into segments
        flat = []
        for reg, part in groups:
            regs = _fill_registers(T, R, tuple(reg))
            for i in range(0, max(len(part), 1), C.SEG_MAX_OPS):
                flat.append((regs, part[i:i + C.SEG_MAX_OPS]))
        segments, cur, cur_ops, cur_f = [], [], 0, 0
        for regs, part in flat:
            f = sum(op.nfloat for op in part)
            if cur and (cur_ops + len(part) > C.SEG_MAX_OPS or len(cur) + 1 > C.SEG_MAX_GROUPS
                        or cur_f + f > C.SEG_MAX_FLOATS):
                segments.append(cur)
                cur, cur_ops, cur_f = [], 0, 0
            cur.append((regs, part))
            cur_ops += len(part)
            cur_f += f
        if cur:
            segments.append(cur)
        # encode
        hdr = [0] * (C.H_SEGOFF + len(segments))
        hdr[C.H_MAGIC], hdr[C.H_T], hdr[C.H_R], hdr[C.H_NSEG] = _i32(C.MAGIC), T, R, len(segments)
        hdr[C.H_L], hdr[C.H_OUTER] = Lp, n_total - T
        for i, g in enumerate(tile_bits):
            hdr[C.H_TBIT + i] = g
        words = list(hdr)
        for si, seg in enumerate(segments):
            words[C.H_SEGOFF + si] = len(words)
            nops = sum(len(p) for _, p in seg)
            seg_floats = sum(op.nfloat for _, p in seg for op in p)
            mat_len = (seg_floats + 3) & ~3
            words += [len(seg), nops, mat_off, mat_len]
            first = 0
            for regs, part in seg:
                rp = list(regs) + [0] * (4 - len(regs))
                words += rp + [len(part), first, 0, 0]
                first += len(part)
            slot = 0
            for regs, part in seg:
                for op in part:
                    pm_t, pm_o = split(op.pmask)
                    cm_t, cm_o = split(op.cmask)
                    cv_t, cv_o = split(op.cval)
                    w0 = op.type | (C.F_GRAD if op.grad else 0) | (C.F_NOUNAPPLY if op.nounapply else 0)
                    tgt = regs.index(op._lt[0]) if op.type == C.T_U1Q else 0
                    if op.type == C.T_U2Q:
                        assert regs[0] == op._lt[1] and regs[1] == op._lt[0]
                    words += [w0, tgt, _i32(pm_t), _i32(cm_t), _i32(cv_t), slot, _i32(pm_o), _i32(cm_o),
                              _i32(cv_o), 0]
                    flags = (C.BF_INVERSE if op.inverse else 0) | (C.BF_PER_SAMPLE if op.pmode == "ps" else 0)
                    coff = 0
                    if op.kind == C.K_CONST:
                        coff = const_off
                        consts.append(op.const)
                        const_off += len(op.const)
                        assert len(op.const) == op.nfloat
                    if op.kind == C.K_EXTERN:
                        st.externs.append((op.rec, mat_off + slot, op.ext_dim, op.inverse))
                    table.append([op.kind, flags, op.poff, mat_off + slot, coff, op.nfloat, op.nparam, 0])
                    slot += op.nfloat
            mat_off += mat_len
        words[C.H_WORDS] = len(words)
        pp = PassProgram()
        pp.T, pp.R, pp.L = T, R, Lp
        pp.prog = np.array(words, dtype=np.int32)
        pp.n_ops, pp.n_groups, pp.n_seg, pp.tile_bits = len(ops), len(flat), len(segments), tile_bits
        st.passes.append(pp)
    st.M = max(mat_off, 4)
    st.table = np.array(table, dtype=np.int32).reshape(-1, C.B_WORDS)
    st.consts = np.concatenate(consts).astype(np.float32) if consts else np.zeros(4, np.float32)
    return st


class CircuitPlan:
    def __init__(self, n):
        self.n = n
        self.stages = []
        self.n_gates = 0

    @property
    def n_passes(self):
        return sum(len(s.passes) for s in self.stages if s.kind == "fused")


def build_plan(n, key, grad_mode, n_total=None, **tile_kw):
    """key: tuple of (name, wires, inverse, pmode, grad, ext_dim) in queue order."""
    plan = CircuitPlan(n)
    plan.n_gates = len(key)
    cur = []

    def close():
        if cur:
            plan.stages.append(build_fused_stage(list(cur), n, n_total, **tile_kw))
            cur.clear()

    for rec, (name, wires, inverse, pmode, grad, ext_dim) in enumerate(key):
        spec = G.SPECS[name]
        if spec.kind == G.NOP:
            continue
        if spec.kind == G.RESET:
            close()
            plan.stages.append(ResetStage(tuple(int(w) for w in wires)))
            continue
        op = lower(n, name, wires, inverse, pmode, grad, rec, ext_dim)
        if op is None:  # dense
            close()
            const = G.qft_matrix_np(len(wires)) if name == "qft" else None
            plan.stages.append(DenseStage(rec, _bits(n, wires), const, inverse, grad_mode))
            continue
        if grad_mode and not spec.unitary:
            # non-unitary in-tile op: its own node, keeps its input instead of un-applying
            close()
            op.nounapply = True
            stg = build_fused_stage([op], n, n_total, **tile_kw)
            stg.save_input = True
            plan.stages.append(stg)
            continue
        cur.append(op)
    close()
    return plan
