"""Host planner: turns a queued gate list into fused passes for the sm_100a kernels.

This replaces the reference's static-mode scheduler (graph/graphs.py:384-509, greedy blocks of
`wires_per_block` qubits multiplied into dense unitaries) with a planner shaped by the memory
system instead: a *pass* is one HBM round trip of the state with a tile of 2^T amplitudes staged
in shared memory; inside a pass, *groups* of gates whose targets fit R register bits are applied
without touching shared memory.  Gates on disjoint qubits commute, so both levels use the same
dependency-respecting greedy packing.  The output is the int32 program of csrc/tqb_program.h.

Lowering picks the cheapest micro-op for every gate (csrc/tqb_program.h, TQB_T_*): X-like gates
become register swaps (PERM), real matrices (ry, h) and rx get 4-instruction-per-amplitude forms,
phase gates (z, s, t, cz, cu1 ...) only touch the amplitudes whose target bit is 1, and
`cnot; diagonal; cnot` sandwiches (the QAOA cost layer) fold into one diagonal op.
"""
import functools
import itertools
import os

import numpy as np

from . import _codes as C
from . import blockplan as B
from . import gates as G

T_SINGLE = 13   # largest state handled as one tile per sample (2^13 amplitudes = 64 KiB)
T_MULTI = 12    # tile bits when the state is streamed through HBM
L_MIN = 5       # contiguous low bits always in the tile (2^5 amplitudes = 256 B runs)
R_MAX = int(os.environ.get("TQB200_R", C.MAX_R))   # register bits per thread (env: kernel experiments)
R_GRAD = int(os.environ.get("TQB200_RGRAD", C.MAX_R))
FAST_FWD = os.environ.get("TQB200_FAST_FWD", "1") != "0"   # forward-only programs: CNOT folding, scaled H, phase ladders
# Scaled rx / ry / one-bit-diagonal forms (TQB_T_RXS / RYS / DIAGS: the gate's cos / sin / d0 factor deferred to the end of
# the pass) are NOT emitted: the un-normalised intermediates grow by up to sqrt(2) per gate, so a deep on-chip pass
# (hundreds of rotations on a <= 13-qubit state) overflowed fp32 while the deferred product underflowed (NaN state), and
# at moderate depth the rounding of the ratio forms drifted the norm (1 + 2.9e-6 after a depth-100 brickwork at n = 14,
# 1 + 5.9e-5 at n = 28) where the plain forms stay at 1 - 4e-7.  The kernel keeps the op types; the planner does not use them.
# Dense 6-qubit tensor-core blocks (blockplan.py) for forward-only programs on streamed states of >= BLOCK_MIN_N local
# qubits: a run of >= BLOCK_MIN_RUN fusable gates becomes a block stage when the schedule averages >= BLOCK_MIN_AVG gates
# per block (one block pass costs about as much as a register-tile pass holding 3 gates: 0.86 ms vs 0.62 + 0.093 k ms at
# n = 28; below ~8 gates per block the deep register-tile passes are cheaper per gate).
BLOCKS = os.environ.get("TQB200_BLOCKS", "1") != "0"
BLOCK_MIN_N = 16
BLOCK_MIN_RUN = 24
BLOCK_MIN_AVG = 8.0
BLOCK_MIN_RUN_KEPT = 6     # ... when the stage may leave its qubit layout behind (no sort passes to pay)
BLOCK_MIN_AVG_KEPT = 3.0
SCALED_ROT = False
SCALED_MAX = 24   # most scaled ops (H with its 1/sqrt(2) deferred) in one pass: intermediates grow by <= 2^12

NFLOAT = {C.T_GEN: 8, C.T_REAL: 4, C.T_RX: 4, C.T_DIAG: 4, C.T_PERM: 0, C.T_U2Q: 32,
          C.T_RXS: 4, C.T_RYS: 4, C.T_HS: 0, C.T_DIAGS: 4}
SCALED = (C.T_RXS, C.T_RYS, C.T_HS, C.T_DIAGS)       # forms whose common factor is deferred to the end of the pass
FOLDS_CNOT = (C.T_RXS, C.T_RYS, C.T_HS, C.T_REAL)    # 1-qubit forms that can absorb a following CNOT on their target
_H = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).astype(np.complex64)
NGRAD = {C.G_NONE: 0, C.G_X: 1, C.G_Y: 1, C.G_DW: 2}


class LOp:
    """A lowered micro-op on global bit positions."""
    __slots__ = ("type", "targets", "pmask", "cmask", "cval", "nfloat", "kind", "inverse", "pmode",
                 "poff", "nparam", "const", "rec", "grad", "gkind", "ngrad", "nounapply", "qmask", "ext_dim", "_lt",
                 "_lc", "_soft", "postx", "scale_const", "rungs", "preh", "unitary")

    def __init__(self):
        self.pmask = self.cmask = self.cval = 0
        self.targets = ()
        self.kind, self.inverse, self.pmode, self.poff, self.nparam = C.K_CONST, False, None, 0, 0
        self.const, self.rec, self.grad, self.nounapply, self.ext_dim = None, -1, False, False, 0
        self.gkind, self.ngrad = C.G_NONE, 0
        self._lc, self._soft, self.postx, self.scale_const = None, (), None, None
        self.rungs = None   # T_LADDER: [(global bit, source LOp)]
        self.preh = False   # T_LADDER: a Hadamard on the control bit is folded in (TQB_F_PREH)
        self.unitary = True

    @property
    def per_sample(self):
        return self.pmode in ("ps", "extps")


def _bits(n, wires):
    return [n - 1 - int(w) for w in wires]


def _applied(mat, inverse):
    m = np.asarray(mat, dtype=np.complex64)
    if inverse:
        m = m.conj().T if m.ndim == 2 else m.conj()
    return np.ascontiguousarray(m)


def _pack_const(m, typ):
    """Applied constant matrix -> the float packing of op type `typ`."""
    if typ == C.T_REAL:
        return np.ascontiguousarray(m.real.astype(np.float32)).reshape(-1)
    return m.view(np.float32).reshape(-1)


def lower(n, name, wires, inverse, pmode, grad, rec, ext_dim=0, fast=False):
    """One queue record -> LOp (fusable) or None for DENSE / RESET / NOP kinds.
    pmode: None | 'sh' | 'ps' | 'ext' / 'extps' (explicit matrix supplied by the caller).
    fast: the program never runs backward -> uncontrolled rx / ry / h / one-bit diagonals take the scaled
    forms of csrc/tqb_program.h (TQB_T_RXS ...)."""
    spec = G.SPECS[name]
    op = LOp()
    op.rec, op.grad, op.inverse, op.pmode = rec, grad, inverse, pmode
    op.unitary = spec.unitary
    bits = _bits(n, wires)
    kind = spec.kind
    const = spec.const
    ext = pmode in ("ext", "extps")
    if name == "qft" and len(wires) <= 2:
        kind = G.U1Q if len(wires) == 1 else G.U2Q
        const = G.qft_matrix_np(len(wires))
    if ext:
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
    if ext:
        n_ctrl = 0
    for b in bits[:n_ctrl]:
        op.cmask |= 1 << b
    op.cval = 0 if spec.neg_ctrl else op.cmask
    rest = bits[n_ctrl:]
    parametric = (not ext) and spec.builder is not None and pmode in ("sh", "ps")
    applied = None if (ext or parametric) else _applied(const, inverse)
    if kind == G.U1Q:
        assert len(rest) == 1, f"{name}: expected one target wire"
        op.targets = (rest[0],)
        if parametric:
            if spec.builder == C.K_RX:
                op.type, op.gkind = (C.T_GEN if op.cmask else (C.T_RXS if fast and SCALED_ROT else C.T_RX)), C.G_X
            elif spec.builder == C.K_RY:
                op.type, op.gkind = (C.T_GEN if op.cmask else (C.T_RYS if fast and SCALED_ROT else C.T_REAL)), C.G_Y
            else:
                op.type, op.gkind = C.T_GEN, C.G_FULL
        elif ext:
            op.type, op.gkind = C.T_GEN, C.G_FULL
        else:
            op.gkind = C.G_NONE
            if np.array_equal(applied, np.array([[0, 1], [1, 0]], dtype=np.complex64)):
                op.type = C.T_PERM
            elif fast and not op.cmask and np.allclose(applied, _H, atol=1e-7):
                op.type, op.scale_const = C.T_HS, complex(np.float32(1 / np.sqrt(2)))
            elif not op.cmask and not np.any(applied.imag):
                op.type = C.T_REAL
            else:
                op.type = C.T_GEN
    elif kind == G.U2Q:
        assert len(rest) == 2, f"{name}: expected two target wires"
        op.type, op.targets, op.gkind = C.T_U2Q, (rest[0], rest[1]), C.G_FULL
    else:
        op.type, op.gkind = C.T_DIAG, C.G_DW
        for b in rest:
            op.pmask |= 1 << b
        phase_like = (spec.builder == C.K_PHASE) if parametric else (applied is not None and applied[0] == 1)
        if fast and SCALED_ROT and len(rest) == 1 and not op.cmask and (parametric or (applied is not None and abs(applied[0]) > 0.5)):
            op.type = C.T_DIAGS   # d0 diag(1, d1/d0): half the amplitudes are touched, d0 joins the pass's scale factor
            if not parametric:
                op.scale_const = complex(applied[0])
        elif phase_like and len(rest) == 1:
            # d0 == 1: only amplitudes whose target bit is set change -> the target doubles as a control
            op.cmask |= op.pmask
            op.cval |= op.pmask
    op.nfloat = NFLOAT[op.type]
    if not grad:
        op.gkind = C.G_NONE
    op.ngrad = op.nfloat if op.gkind == C.G_FULL else NGRAD[op.gkind]
    if not ext:
        if parametric:
            op.kind, op.nparam = spec.builder, spec.n_params
        else:
            op.kind, op.const = C.K_CONST, _pack_const(applied, op.type)
            if op.type in (C.T_PERM, C.T_HS):
                op.const = np.zeros(0, np.float32)
            elif op.type == C.T_DIAGS:
                rel = np.complex64(applied[1] / applied[0])
                op.const = np.array([rel.real, rel.imag, applied[0].real, applied[0].imag], dtype=np.float32)
    tm = 0
    for t in op.targets:
        tm |= 1 << t
    op.qmask = tm | op.pmask | op.cmask
    assert bin(op.qmask).count("1") == len(set(bits)), f"{name}: repeated wires {wires}"
    return op


def fold_cnot_sandwiches(lops):
    """cnot(c,t); D; cnot(c,t) with D an uncontrolled parity-diagonal op acting on t (not on c)
    == the same diagonal with c added to its parity mask.  Removes two of the three ops of every
    `cx, rz, cx` edge term of a QAOA / ZZ-feature-map layer."""
    out, i = [], 0
    while i < len(lops):
        a = lops[i]
        if (i + 2 < len(lops) and a.type == C.T_PERM and bin(a.cmask).count("1") == 1 and a.cval == a.cmask):
            d, b = lops[i + 1], lops[i + 2]
            t, cbit = a.targets[0], a.cmask
            if (b.type == C.T_PERM and b.cmask == a.cmask and b.cval == a.cval and b.targets == a.targets
                    and d.type in (C.T_DIAG, C.T_DIAGS) and d.cmask == 0 and (d.pmask >> t) & 1 and not d.pmask & cbit
                    and not d.nounapply):
                if d.type == C.T_DIAGS:   # two parity bits now: back to the (d0, d1) form
                    d.type, d.scale_const = C.T_DIAG, None
                    if d.kind == C.K_CONST:
                        rel, d0 = complex(d.const[0], d.const[1]), complex(d.const[2], d.const[3])
                        d1 = rel * d0
                        d.const = np.array([d0.real, d0.imag, d1.real, d1.imag], dtype=np.float32)
                d.pmask ^= cbit
                d.qmask |= cbit
                out.append(d)
                i += 3
                continue
        out.append(a)
        i += 1
    return out


LADDER_MIN = 3   # shortest run of controlled phases worth a ladder op


def _is_rung(op):
    """cu1 / cphase-like: phase e^{i theta} where two bits are both set, nothing else; shared or constant angle."""
    if op.type != C.T_DIAG or op.grad or op.nounapply or op.pmode not in (None, "sh"):
        return False
    if bin(op.cmask).count("1") != 2 or op.cval != op.cmask or (op.pmask & ~op.cmask) or bin(op.pmask).count("1") != 1:
        return False
    if op.kind == C.K_CONST:
        return op.const is not None and op.const[0] == 1 and op.const[1] == 0
    return op.kind == C.K_PHASE


def fold_phase_ladders(lops):
    """Forward-only programs: a run of >= LADDER_MIN consecutive controlled phases that share one qubit c and hit
    distinct partners (the rungs H_c, cu1(c+1, c), cu1(c+2, c), ... of a QFT) becomes ONE TQB_T_LADDER op:
    where bit c is set, the amplitude is multiplied by the product of the phases of the partner bits that are set.
    One op record and ~200 instructions per thread instead of one masked diagonal op per rung."""
    out, i = [], 0
    while i < len(lops):
        a = lops[i]
        if _is_rung(a) and i + 1 < len(lops) and _is_rung(lops[i + 1]) and (a.cmask & lops[i + 1].cmask):
            common = a.cmask & lops[i + 1].cmask
            if bin(common).count("1") == 1:
                run, seen, k = [], 0, i
                while k < len(lops) and _is_rung(lops[k]) and (lops[k].cmask & common):
                    partner = lops[k].cmask & ~common
                    if partner == 0 or (partner & seen):
                        break
                    seen |= partner
                    run.append(lops[k])
                    k += 1
                if len(run) >= LADDER_MIN:
                    lad = LOp()
                    lad.type, lad.rec, lad.pmode = C.T_LADDER, run[0].rec, None
                    lad.cmask = lad.cval = common
                    lad.pmask = seen
                    lad.qmask = common | seen
                    lad.rungs = [((r.cmask & ~common).bit_length() - 1, r) for r in run]
                    lad.nfloat = 0   # sized per pass (2 floats per tile-local and per outer bit position)
                    lad.kind = C.K_CONST
                    out.append(lad)
                    i = k
                    continue
        out.append(a)
        i += 1
    return out


def fold_h_into_ladders(lops):
    """`H on c; ladder controlled by c` (one rung set of a QFT) -> the ladder with TQB_F_PREH: one sweep over the staged tile
    does (a, b) -> (a + b, (a - b) * phase) on the pairs of bit c instead of a register group for H plus a ladder group."""
    out, i = [], 0
    while i < len(lops):
        a = lops[i]
        if (a.type == C.T_HS and i + 1 < len(lops) and lops[i + 1].type == C.T_LADDER and not lops[i + 1].preh
                and lops[i + 1].cmask == 1 << a.targets[0] and not (lops[i + 1].pmask >> a.targets[0]) & 1):
            lad = lops[i + 1]
            lad.preh, lad.targets, lad.scale_const = True, a.targets, a.scale_const
            out.append(lad)
            i += 2
            continue
        out.append(a)
        i += 1
    return out


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
        reg, blocked, part, rest, soft, ladder_group, closed = [], 0, [], [], [], False, False
        for op in remaining:
            ok = not (op.qmask & blocked)
            if ok and (op.type == C.T_LADDER) != ladder_group and part:
                ok = False   # ladders run in groups of their own (applied on the staged tile, no register bits)
            if ok and part and (closed or op.preh):
                ok = False   # a ladder with a folded Hadamard pairs amplitudes of different threads: one per group (barrier)
            if ok and not part:
                ladder_group = op.type == C.T_LADDER
            if ok:
                if op.type == C.T_U2Q:
                    tm, tl = op._lt
                    if not reg and R >= 2:
                        reg.extend([tl, tm])
                    elif not (len(reg) >= 2 and reg[0] == tl and reg[1] == tm):
                        ok = False
                elif op._lt:
                    t = op._lt[0]
                    if t not in reg:
                        if len(reg) < R:
                            reg.append(t)
                        else:
                            ok = False
            if ok:
                soft.extend(op._soft)
                part.append(op)
                closed = closed or op.preh
            else:
                blocked |= op.qmask
                rest.append(op)
        for b in soft:   # bits that are cheaper as register bits but need not be (fast programs): spare slots only
            if b not in reg and len(reg) < R:
                reg.append(b)
        groups.append((reg, part))
        remaining = rest
    return groups


SCALED_MIN = 3   # fewer scaled ops than this in a pass: the deferred factor costs more than the forms save


def _unscale(op):
    """Scaled form -> the plain form of the same gate (a pass with only one or two of them pays more for applying
    the deferred factor at write-back, 4 FP instructions per amplitude, than the forms save)."""
    if op.type == C.T_RXS:
        op.type = C.T_RX
    elif op.type == C.T_RYS:
        op.type = C.T_REAL
    elif op.type == C.T_HS:
        op.type, op.const = C.T_REAL, _pack_const(_H, C.T_REAL)
    elif op.type == C.T_DIAGS:
        op.type = C.T_DIAG
        if op.kind == C.K_CONST:
            rel, d0 = complex(op.const[0], op.const[1]), complex(op.const[2], op.const[3])
            d1 = rel * d0
            op.const = np.array([d0.real, d0.imag, d1.real, d1.imag], dtype=np.float32)
    op.scale_const = None
    op.nfloat = NFLOAT[op.type]


def _fold_cnots(part, regs):
    """Fast programs: `G on t; cnot(c -> t)` with c and t both register bits of the group and G an uncontrolled
    one-qubit op on t becomes G with TQB_F_POSTX (the kernel exchanges the pair's outputs where c is set, which
    is register renaming).  G may be moved up to the CNOT: everything between acts on other bits."""
    regset = set(regs)
    out = []
    for op in part:
        if (op.type == C.T_PERM and op._lc is not None and op._lc in regset and op._lt[0] in regset
                and op._lc != op._lt[0]):
            tg = op.targets[0]
            for k in range(len(out) - 1, -1, -1):
                g = out[k]
                if (g.qmask >> tg) & 1:
                    one_bit = (g.type in FOLDS_CNOT and g.targets == (tg,) and not g.cmask) or \
                        (g.type == C.T_DIAGS and g.pmask == 1 << tg)
                    if one_bit and g.postx is None:
                        out.pop(k)
                        g.postx = op._lc
                        out.append(g)
                        op = None
                    break
        if op is not None:
            out.append(op)
    return out


def _swz(idx, pair):
    # csrc/tqb_core.cuh swz(): pair layout for streamed states, 4-bit xor for single-tile states
    return idx ^ (((idx >> 4) & 7) << 1) if pair else idx ^ ((idx >> 4) & 15)


@functools.lru_cache(maxsize=4096)
def _fill_registers(T, R, reg, pair=False):
    """Complete `reg` (tuple of tile-local target bits) to R distinct bits, choosing the filler bits
    that minimise shared-memory bank conflicts of the group's 64-bit accesses."""
    reg = list(reg)
    free = [b for b in range(T) if b not in reg]
    missing = R - len(reg)
    if missing == 0:
        return tuple(reg)

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
                k = _swz(w, pair) & 15
                banks[k] = banks.get(k, 0) + 1
            if banks:
                worst = max(worst, max(banks.values()))
        return worst

    best, best_deg = None, None
    for combo in itertools.combinations(free, missing):
        d = degree(reg + list(combo))
        if best is None or d < best_deg:
            best, best_deg = combo, d
            if d == 1:
                break
    return tuple(reg + list(best))


class PassProgram:
    __slots__ = ("T", "R", "L", "prog", "n_ops", "n_groups", "n_seg", "tile_bits", "desc", "_desc_c")


class FusedStage:
    """One autograd node: a list of passes sharing one matrix / gradient row layout."""
    kind = "fused"

    def __init__(self):
        self.passes = []
        self.table = None      # int32 [nrec, B_WORDS], shared-row records first
        self.n_sh_rec = 0
        self.consts = None     # float32
        self.M_sh = self.M_ps = self.G = 0   # floats per shared matrix row / per-sample matrix row / gradient row
        self.P_sh = self.P_ps = 0
        self.sh_recs, self.ps_recs = [], []   # queue indices in parameter order
        self.externs = []      # (rec, moff, dim, inverse, per_sample_row, goff)
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


def _encode_op(op, regs, R, loc, oloc):
    """Op record for a group whose register bit j is tile-local bit regs[j]."""
    regidx = {b: j for j, b in enumerate(regs[:R])}
    cm_t = cv_t = pm_t = cm_o = cv_o = pm_o = 0
    reg_ctrl, reg_par = [], []

    def walk(mask):
        p = 0
        while mask >> p:
            if (mask >> p) & 1:
                yield p
            p += 1

    for p in walk(op.cmask):
        want = (op.cval >> p) & 1
        if p in loc:
            lb = loc[p]
            if lb in regidx:
                reg_ctrl.append((regidx[lb], want))
            else:
                cm_t |= 1 << lb
                cv_t |= want << lb
        else:
            cm_o |= 1 << oloc[p]
            cv_o |= want << oloc[p]
    for p in walk(op.pmask):
        if p in loc:
            lb = loc[p]
            if lb in regidx:
                reg_par.append(regidx[lb])
            else:
                pm_t |= 1 << lb
        else:
            pm_o |= 1 << oloc[p]
    emask = epar = 0
    for e in range(16):
        if all(((e >> j) & 1) == want for j, want in reg_ctrl):
            emask |= 1 << e
        if sum((e >> j) & 1 for j in reg_par) & 1:
            epar |= 1 << e
    flags = 0
    if cm_t:
        flags |= C.F_TCTRL
    if emask != 0xFFFF:
        flags |= C.F_RCTRL
    if op.per_sample:
        flags |= C.F_PS
    if op.nounapply:
        flags |= C.F_NOUNAPPLY
    if cm_o:
        flags |= C.F_OCTRL
    j = 0
    if op.type in (C.T_DIAG, C.T_DIAGS) and len(reg_par) == 1:
        flags |= C.F_DJ
        j = reg_par[0]
    if op.type == C.T_U2Q:
        assert regs[0] == op._lt[1] and regs[1] == op._lt[0]
    elif op._lt and op.type != C.T_LADDER:
        j = regidx[op._lt[0]]
    if op.type in (C.T_REAL, C.T_RX):
        assert not (flags & (C.F_TCTRL | C.F_RCTRL)), "REAL / RX micro-ops cannot be controlled"
    w0 = (op.type << 2) | j | flags | (op.gkind << 16)
    if op.type == C.T_LADDER and op.preh:
        w0 |= C.F_PREH | (op._lt[0] << 22)
    if op.postx is not None:
        assert op.postx in regidx and regidx[op.postx] != j and (op.type != C.T_DIAGS or flags & C.F_DJ)
        w0 |= C.F_POSTX | (regidx[op.postx] << 20)
    w = [0] * C.OP_WORDS
    w[C.OP_W0], w[C.OP_EMASK] = w0, emask | (epar << 16)
    w[C.OP_CM_T], w[C.OP_CV_T], w[C.OP_PM_T] = cm_t, cv_t, pm_t
    w[C.OP_CM_O], w[C.OP_CV_O], w[C.OP_PM_O] = cm_o, cv_o, pm_o
    return w


def build_fused_stage(lops, n_local, n_total=None, t_single=None, t_multi=None, l_min=None, r=None, fast=False):
    """lops: LOps in program order (targets < n_local; masks may use bits up to n_total).
    fast: forward-only program (scaled op forms, CNOT folding)."""
    n_total = n_local if n_total is None else n_total
    t_single = T_SINGLE if t_single is None else t_single
    t_multi = T_MULTI if t_multi is None else t_multi
    l_min = L_MIN if l_min is None else l_min
    st = FusedStage()
    st.n_ops = len(lops)
    # parameter slots in queue order
    for top in lops:
        for op in ([r for _, r in top.rungs] if top.rungs else [top]):
            if op.pmode == "sh":
                op.poff = st.P_sh
                st.P_sh += op.nparam
                st.sh_recs.append(op.rec)
            elif op.pmode == "ps":
                op.poff = st.P_ps
                st.P_ps += op.nparam
                st.ps_recs.append(op.rec)
            st.any_grad |= op.gkind != C.G_NONE
    # passes
    t_single = max(t_single, min(t_multi, n_local))
    if n_local <= t_single:
        raw = [(set(), list(lops))]
        T, L = n_local, n_local
    else:
        T, L = t_multi, l_min
        raw = _pack_passes(lops, L, T - L)
    R = min(R_MAX if r is None else r, T)
    consts, table_sh, table_ps = [], [], []
    const_off = 0
    sh_off = ps_off = g_off = 0
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
        n_scaled = sum(op.type in SCALED or op.preh for op in ops)
        if fast and (n_scaled < SCALED_MIN or n_scaled > SCALED_MAX):
            plain = []
            for op in ops:
                if op.type in SCALED:
                    _unscale(op)
                elif op.preh and n_scaled > SCALED_MAX:
                    # un-fold the Hadamard of a ladder: a plain H in front, the ladder without TQB_F_PREH
                    h = lower(n_total, "hadamard", [n_total - 1 - op.targets[0]], False, None, False, op.rec)
                    op.preh, op.targets, op.scale_const = False, (), None
                    plain.append(h)
                plain.append(op)
            ops = plain
        for op in ops:
            op._lt = tuple(loc[t] for t in op.targets)
            op._lc, op._soft = None, ()
            if fast and op.type == C.T_PERM and bin(op.cmask).count("1") == 1 and op.cval == op.cmask:
                cb = op.cmask.bit_length() - 1
                if cb in loc:
                    op._lc, op._soft = loc[cb], (loc[cb],)
            elif fast and op.type == C.T_DIAGS and (op.pmask.bit_length() - 1) in loc:
                op._soft = (loc[op.pmask.bit_length() - 1],)
            if op.type == C.T_LADDER:
                op.nfloat = (2 * (T + len(outer)) + 3) & ~3
        groups = _pack_groups(ops, R)
        sh_rec0, ps_rec0 = len(table_sh), len(table_ps)
        # split oversize groups, then cut into segments
        flat = []
        for reg, part in groups:
            regs = _fill_registers(T, R, tuple(reg), n_local > t_single)
            if fast:
                part = _fold_cnots(part, regs)
            # a chunk of a group must fit one segment's staging capacity on its own
            chunk, use = [], [0, 0, 0, 0]
            for op in part:
                need = [1, 0 if op.per_sample else op.nfloat, op.nfloat if op.per_sample else 0, op.ngrad]
                if chunk and (use[0] + 1 > C.SEG_MAX_OPS or use[1] + need[1] > C.SEG_MAX_SH
                              or use[2] + need[2] > C.SEG_MAX_PS or use[3] + need[3] > C.SEG_MAX_G):
                    flat.append((regs, chunk))
                    chunk, use = [], [0, 0, 0, 0]
                chunk.append(op)
                use = [a + b for a, b in zip(use, need)]
            flat.append((regs, chunk))
        segments, cur, tot = [], [], [0, 0, 0, 0]
        for regs, part in flat:
            need = [len(part), sum(op.nfloat for op in part if not op.per_sample),
                    sum(op.nfloat for op in part if op.per_sample), sum(op.ngrad for op in part)]
            if cur and (tot[0] + need[0] > C.SEG_MAX_OPS or len(cur) + 1 > C.SEG_MAX_GROUPS
                        or tot[1] + need[1] > C.SEG_MAX_SH or tot[2] + need[2] > C.SEG_MAX_PS
                        or tot[3] + need[3] > C.SEG_MAX_G):
                segments.append(cur)
                cur, tot = [], [0, 0, 0, 0]
            cur.append((regs, part))
            tot = [a + b for a, b in zip(tot, need)]
            assert tot[1] <= C.SEG_MAX_SH and tot[2] <= C.SEG_MAX_PS and tot[3] <= C.SEG_MAX_G
        if cur:
            segments.append(cur)
        # encode
        hdr = [0] * (C.H_SEGOFF + len(segments))
        hdr[C.H_MAGIC], hdr[C.H_T], hdr[C.H_R], hdr[C.H_NSEG] = _i32(C.MAGIC), T, R, len(segments)
        hdr[C.H_L], hdr[C.H_OUTER] = Lp, n_total - T
        hdr[C.H_SCALE_SH] = hdr[C.H_SCALE_PS] = -1
        for i, g in enumerate(tile_bits):
            hdr[C.H_TBIT + i] = g
        words = list(hdr)
        sh_max = ps_max = g_max = 0
        for si, seg in enumerate(segments):
            words[C.H_SEGOFF + si] = len(words)
            nops = sum(len(p) for _, p in seg)
            seg_ops = [op for _, p in seg for op in p]
            sh_len = sum(op.nfloat for op in seg_ops if not op.per_sample)
            ps_len = sum(op.nfloat for op in seg_ops if op.per_sample)
            g_len = (sum(op.ngrad for op in seg_ops) + 3) & ~3
            sh_max, ps_max, g_max = max(sh_max, sh_len), max(ps_max, ps_len), max(g_max, g_len)
            words += [len(seg), nops, sh_off, sh_len, ps_off, ps_len, g_off, g_len]
            first = 0
            n_ladders = sum(op.type == C.T_LADDER for op in seg_ops)
            for regs, part in seg:
                rp = list(regs) + [0] * (4 - len(regs))
                is_lad = 1 if part and part[0].type == C.T_LADDER else 0
                words += rp + [len(part), first, n_ladders if first == 0 else 0, is_lad]
                first += len(part)
            sh_slot = ps_slot = g_slot = 0
            for regs, part in seg:
                for op in part:
                    w = _encode_op(op, () if op.type == C.T_LADDER else regs, R, loc, oloc)
                    if op.per_sample:
                        w[C.OP_SLOT], moff = ps_slot, ps_off + ps_slot
                        ps_slot += op.nfloat
                    else:
                        w[C.OP_SLOT], moff = sh_slot, sh_off + sh_slot
                        sh_slot += op.nfloat
                    w[C.OP_GSLOT], goff = g_slot, g_off + g_slot
                    g_slot += op.ngrad
                    words += [_i32(v) for v in w]
                    if op.nfloat == 0:
                        continue
                    if op.type == C.T_LADDER:
                        # one 2-float record per rung at its bit's table position, identity (1, 0) in the gaps
                        # (records are evaluated concurrently: no two may write the same floats)
                        rung_at = {(loc[bit] if bit in loc else T + oloc[bit]): r for bit, r in op.rungs}
                        npos, pos = op.nfloat // 2, 0
                        while pos < npos:
                            r = rung_at.get(pos)
                            if r is None:
                                end = pos
                                while end < npos and end not in rung_at:
                                    end += 1
                                consts.append(np.tile(np.array([1.0, 0.0], dtype=np.float32), end - pos))
                                table_sh.append([C.K_CONST, 0, 0, moff + 2 * pos, const_off, 2 * (end - pos), 0, C.T_LADDER,
                                                 C.G_NONE, 0])
                                const_off += 2 * (end - pos)
                                pos = end
                                continue
                            if r.kind == C.K_CONST:
                                consts.append(np.asarray(r.const[2:4], dtype=np.float32))
                                table_sh.append([C.K_CONST, 0, 0, moff + 2 * pos, const_off, 2, 0, C.T_LADDER, C.G_NONE, 0])
                                const_off += 2
                            else:
                                table_sh.append([r.kind, C.BF_INVERSE if r.inverse else 0, r.poff, moff + 2 * pos, 0, 2,
                                                 r.nparam, C.T_LADDER, C.G_NONE, 0])
                            pos += 1
                        continue
                    flags = (C.BF_INVERSE if op.inverse else 0) | (C.BF_PER_SAMPLE if op.pmode == "ps" else 0)
                    coff = 0
                    if op.kind == C.K_CONST:
                        coff = const_off
                        consts.append(op.const)
                        const_off += len(op.const)
                        assert len(op.const) == op.nfloat
                    if op.kind == C.K_EXTERN:
                        st.externs.append((op.rec, moff, op.ext_dim, op.inverse, op.per_sample, goff))
                    rec = [op.kind, flags, op.poff, moff, coff, op.nfloat, op.nparam, op.type, op.gkind, goff]
                    (table_ps if op.per_sample else table_sh).append(rec)
            sh_off += sh_len
            ps_off += ps_len
            g_off += g_len
        # deferred scale factor of the pass (scaled op forms): constant part folded here, parametric part by the
        # matrix builder (K_SCALE records scan this pass's records)
        scaled = [op for op in ops if op.type in SCALED or op.scale_const is not None]
        if scaled:
            pre = complex(1.0)
            for op in scaled:
                if op.scale_const is not None:
                    pre *= op.scale_const
            if pre != 1.0 or any(op.scale_const is None and not op.per_sample for op in scaled):
                consts.append(np.array([pre.real, pre.imag], dtype=np.float32))
                table_sh.append([C.K_SCALE, 0, const_off, sh_off, sh_rec0, len(table_sh) - sh_rec0, 0, 0, C.G_NONE, 0])
                const_off += 2
                words[C.H_SCALE_SH] = sh_off
                sh_off += 4
            if any(op.per_sample for op in scaled):
                consts.append(np.array([1.0, 0.0], dtype=np.float32))
                # COFF is local to the per-sample table here; made absolute once the shared table is complete
                table_ps.append([C.K_SCALE, C.BF_PER_SAMPLE, const_off, ps_off, ps_rec0, len(table_ps) - ps_rec0, 0, 0,
                                 C.G_NONE, 0])
                const_off += 2
                words[C.H_SCALE_PS] = ps_off
                ps_off += 4
        words[C.H_WORDS] = len(words)
        words[C.H_SH_MAX], words[C.H_PS_MAX], words[C.H_G_MAX] = sh_max, ps_max, g_max
        pp = PassProgram()
        pp.T, pp.R, pp.L = T, R, Lp
        pp.prog = np.array(words, dtype=np.int32)
        pp.n_ops, pp.n_groups, pp.n_seg, pp.tile_bits = len(ops), len(flat), len(segments), tile_bits
        pp.desc = [T, R, Lp, len(segments), sh_max, ps_max, g_max, 0]
        pp._desc_c = None
        st.passes.append(pp)
    st.M_sh, st.M_ps, st.G = max(sh_off, 4), max(ps_off, 4), max(g_off, 4)
    st.n_sh_rec = len(table_sh)
    for rec in table_ps:
        if rec[C.B_KIND] == C.K_SCALE:
            rec[C.B_COFF] += len(table_sh)
    st.table = np.array(table_sh + table_ps, dtype=np.int32).reshape(-1, C.B_WORDS)
    st.consts = np.concatenate(consts + [np.zeros(4, np.float32)]).astype(np.float32)
    return st


def _split_block_runs(ops, n_local, leave_last=False):
    """Forward-only program -> [("blocks", BlockStage) | ("fused", [LOp])]: maximal runs of block-fusable ops that the
    block scheduler packs densely enough go to the tensor-core block kernel, everything else (ladders, per-sample
    parameters, explicit matrices, sparse runs) stays with the register-tile pass kernel."""
    out, i = [], 0

    def fused(part):
        if out and out[-1][0] == "fused":
            out[-1][1].extend(part)
        else:
            out.append(("fused", list(part)))

    while i < len(ops):
        k = i
        if B.fusable(ops[i], n_local):
            while k < len(ops) and B.fusable(ops[k], n_local):
                k += 1
            run = ops[i:k]
            # the last run of a plan whose caller keeps the layout (sharded device) pays no way home: blocks win against
            # register-tile passes much earlier there (a register-tile pass over high free bits writes 256-byte rows at
            # large strides, the slow pattern; 16 X gates of a state preparation at 31 local qubits: 3 passes, 39 ms)
            free = leave_last and k == len(ops)
            stage = (B.build_block_stage(run, n_local, BLOCK_MIN_AVG_KEPT if free else BLOCK_MIN_AVG, come_home=not free)
                     if len(run) >= (BLOCK_MIN_RUN_KEPT if free else BLOCK_MIN_RUN) else None)
            if stage is not None:
                out.append(("blocks", stage))
            else:
                fused(run)
        else:
            while k < len(ops) and not B.fusable(ops[k], n_local):
                k += 1
            fused(ops[i:k])
        i = k
    return out


class CircuitPlan:
    def __init__(self, n):
        self.n = n
        self.stages = []
        self.n_gates = 0
        self.final_pos = None   # leave_layout plans: logical -> physical bit of the local qubits after the last stage

    @property
    def n_passes(self):
        return sum(len(s.passes) if s.kind == "fused" else s.n_blocks + len(s.quads) for s in self.stages if s.kind in ("fused", "blocks"))


def build_plan(n, key, grad_mode, n_total=None, leave_layout=False, **tile_kw):
    """key: tuple of (name, wires, inverse, pmode, grad, ext_dim) in queue order.

    Sharded statevector (n_total > n): this rank holds the n low index bits, `wires` count from the
    most significant of all n_total bits (wire w = bit n_total-1-w); wires 0 .. n_total-n-1 live in
    the rank index and may only be controls or parity bits of diagonal gates (K11: no traffic)."""
    plan = CircuitPlan(n)
    nb = n if n_total is None else n_total
    plan.n_gates = len(key)
    cur = []
    if grad_mode and "r" not in tile_kw:
        tile_kw = dict(tile_kw, r=R_GRAD)
    fast = FAST_FWD and not grad_mode
    if fast:
        tile_kw = dict(tile_kw, fast=True)

    def close(last=False):
        if cur:
            ops = fold_cnot_sandwiches(cur)
            # block stages are cut out of the UNFOLDED program (controlled phases stay 2-qubit diagonals there: inside a
            # block, or many at once in a quadratic-phase pass); what is left gets the ladder folds of the register-tile kernel
            for kind, part in (_split_block_runs(ops, n, leave_last=last and leave_layout)
                               if (fast and BLOCKS and n >= BLOCK_MIN_N) else [("fused", ops)]):
                if kind == "blocks":
                    plan.stages.append(part)
                else:
                    if fast:
                        part = fold_h_into_ladders(fold_phase_ladders(part))
                    plan.stages.append(build_fused_stage(part, n, n_total, **tile_kw))
            cur.clear()

    for rec, (name, wires, inverse, pmode, grad, ext_dim) in enumerate(key):
        spec = G.SPECS[name]
        if spec.kind == G.NOP:
            continue
        if spec.kind == G.RESET:
            close()
            plan.stages.append(ResetStage(tuple(int(w) for w in wires)))
            continue
        op = lower(nb, name, wires, inverse, pmode, grad, rec, ext_dim, fast=fast)
        if op is None:  # dense
            close()
            const = G.qft_matrix_np(len(wires)) if name == "qft" else None
            bits = _bits(nb, wires)
            assert max(bits) < n, f"{name}: dense gates need all their wires on this rank"
            plan.stages.append(DenseStage(rec, bits, const, inverse, grad_mode))
            continue
        assert all(t < n for t in op.targets), f"{name}: target wire lives in the rank index (swap it in first)"
        if grad_mode and not spec.unitary:
            # non-unitary in-tile op: its own node, keeps its input instead of un-applying
            close()
            op.nounapply = True
            stg = build_fused_stage([op], n, n_total, **tile_kw)
            stg.save_input = True
            plan.stages.append(stg)
            continue
        cur.append(op)
    close(last=True)
    if plan.stages and plan.stages[-1].kind == "blocks":
        plan.final_pos = plan.stages[-1].final_pos     # (only set when leave_layout asked for it)
    return plan
