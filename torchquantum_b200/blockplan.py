"""Block scheduler for forward-only programs on streamed states: packs runs of gates on <= 6 qubits into dense 6-qubit
blocks, each of which costs ONE tensor-core pass over the state (csrc/tqb_block_umma.cuh) whatever it holds.

This is the B200 counterpart of the reference's static-mode scheduler (torchquantum/graph/graphs.py:384-509: greedy blocks
of `wires_per_block` qubits, :713-861 block unitaries on the host).  Differences that the memory system dictates:

* the block kernel keeps index bits 0-5 inside its 128-byte rows, so a block can only act on physical bits >= 6.  The
  six logical qubits that live in bits 0-5 are *parked*.  A block may write its tile back with the roles of its own six
  bits and bits 0-5 exchanged (TQB_KBF_TRANSPOSE) at no cost, which un-parks them.  The scheduler only ever uses two
  layouts: identity (logical qubits 0-5 parked) and "K" (the six qubits K of the last transposing block parked, logical
  qubit b < 6 at physical bit K[b]); the block that acts on logical qubits 0-5 transposes back;
* a block's six physical bits must form <= 3 contiguous runs (rank-5 TMA tensor map: 128-byte row, row index, 3 runs).

Block choice is greedy with connected growth: a seed gate on the dependency frontier, then every later gate that touches
the block and still fits, then a free fill; the candidate absorbing the most gates wins.

Two write-back disciplines (the same block choice logic serves both):

* in place (`schedule`): every tile goes back where it came from; the qubit layout only ever changes through the
  transposing blocks described above.  Blocks whose six bits are all high write 512-byte rows at strides >= 4 MiB, which
  the memory system takes at 0.6x the speed of any other placement (profiles/r02_block6_kpos_sweep.txt: 1.1-1.37 ms
  against 0.80-0.85 ms per pass at n = 28; the gathering READS of the same tiles run at full speed);
* rotating (`schedule_rotating`, TQB_KBF_ROTATE): every block reads its tiles from wherever its qubits are and writes
  each tile as one contiguous 32 KiB piece into a second buffer, so the block's qubits land in index bits 6-11 (or 0-5
  when it also transposes) and all other bits >= 6 move up in their previous relative order.  Any six un-parked qubits
  are a block (bits beyond three contiguous runs are enumerated by the kernel); the scheduler tracks the logical ->
  physical map and ends the stage with <= ceil((n-6)/6) gate-free passes that sort the qubits home.
`build_block_stage` plans both and keeps the cheaper one by a per-placement cost model taken from the sweep."""
import os

import numpy as np

from . import _codes as C

K_BLOCK = 6
LOW = (1 << K_BLOCK) - 1   # logical qubits 0..5 (parked in the identity layout)
MAX_RUNS = 3
_H = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).astype(np.complex64)

FUSABLE_TYPES = (C.T_GEN, C.T_REAL, C.T_RX, C.T_DIAG, C.T_PERM, C.T_U2Q, C.T_HS)


def fusable(op, n_local):
    """Can this lowered op be multiplied into a shared block unitary?"""
    if op.type not in FUSABLE_TYPES or op.pmode not in (None, "sh") or op.kind == C.K_EXTERN or not op.unitary:
        return False
    if op.postx is not None or op.nounapply:
        return False
    local = op.qmask & ((1 << n_local) - 1)
    return bin(local).count("1") <= K_BLOCK


def _runs(bits):
    r, prev = 0, None
    for b in bits:
        if prev is None or b != prev + 1:
            r += 1
        prev = b
    return r


def pad_kpos(phys, n_local):
    """phys: physical bits (all in [6, n_local)) a block must hold -> ascending list of exactly 6 bits with <= MAX_RUNS
    runs, or None.  Pad bits first close the smallest gaps between runs, then extend a run."""
    bits = sorted(phys)
    if len(bits) > K_BLOCK or (bits and (bits[0] < K_BLOCK or bits[-1] >= n_local)):
        return None
    have = set(bits)
    budget = K_BLOCK - len(bits)
    while budget > 0:
        bits = sorted(have)
        gaps = [(bits[i + 1] - bits[i] - 1, bits[i]) for i in range(len(bits) - 1) if bits[i + 1] - bits[i] > 1]
        gaps = [g for g in gaps if g[0] <= budget]
        if gaps and _runs(bits) > 1:
            gap, lo = min(gaps)
            for b in range(lo + 1, lo + 1 + gap):
                have.add(b)
            budget -= gap
            continue
        # extend any run by one bit (up first, then down); an empty set starts at bit 6
        if not bits:
            have.add(K_BLOCK)
            budget -= 1
            continue
        grown = False
        for b in bits:
            for cand in (b + 1, b - 1):
                if K_BLOCK <= cand < n_local and cand not in have:
                    have.add(cand)
                    budget -= 1
                    grown = True
                    break
            if grown:
                break
        if not grown:
            return None
    bits = sorted(have)
    return bits if _runs(bits) <= MAX_RUNS else None


class _Sched:
    """masks: logical local-qubit mask of every gate, program order."""

    def __init__(self, masks, n_local, rotating=False):
        self.masks, self.n = masks, n_local
        self.full = (1 << n_local) - 1
        self.parked = LOW          # logical qubits in physical bits 0-5
        self.K = None              # K layout: sorted logical qubits parked; logical b < 6 sits at physical K[b]
        self.rotating = rotating   # rotating blocks: any six un-parked qubits can be one block
        self._feas = {}

    def phys(self, l):
        if self.K is None:
            return l
        return self.K[l] if l < K_BLOCK else l

    def feasible(self, mask):
        """the logical set can be one block in the current layout"""
        if self.rotating:
            return not (mask & self.parked) and bin(mask).count("1") <= K_BLOCK
        key = (mask, None if self.K is None else tuple(self.K))
        hit = self._feas.get(key)
        if hit is None:
            ls = [b for b in range(self.n) if (mask >> b) & 1]
            hit = len(ls) <= K_BLOCK and not (mask & self.parked) and pad_kpos([self.phys(l) for l in ls], self.n) is not None
            self._feas[key] = hit
        return hit

    def first_fit(self, remaining, forbidden, chosen):
        masks, full = self.masks, self.full
        blocked, taken = forbidden, []
        for idx in remaining:
            q = masks[idx]
            if q & blocked:
                blocked |= q
            else:
                new = q & ~chosen
                if not new:
                    taken.append(idx)
                elif self.feasible(chosen | new):
                    chosen |= new
                    taken.append(idx)
                else:
                    blocked |= q
            if blocked & ~forbidden == full & ~forbidden:
                break
        return chosen, taken

    def grow(self, remaining, forbidden, seed):
        masks, full = self.masks, self.full
        chosen, blocked = seed, forbidden
        for idx in remaining:
            q = masks[idx]
            if q & blocked:
                blocked |= q
            elif q & chosen:
                new = q & ~chosen
                if not new or self.feasible(chosen | new):
                    chosen |= new
                else:
                    blocked |= q
            else:
                blocked |= q
            if blocked & ~forbidden == full & ~forbidden:
                break
        return self.first_fit(remaining, forbidden, chosen)

    def best_block(self, remaining, forbidden):
        masks, full = self.masks, self.full
        blocked, seeds = forbidden, []
        for idx in remaining:
            q = masks[idx]
            if q & blocked:
                blocked |= q
            else:
                if self.feasible(q):
                    seeds.append(q)
                blocked |= q
            if blocked == full:
                break
        best = None
        for s in seeds:
            ch, tk = self.grow(remaining, forbidden, s)
            if best is None or len(tk) > len(best[1]):
                best = (ch, tk)
        return best if best is not None else (0, [])

    def exact_low(self, remaining):
        """gates that fit a block on exactly the logical qubits 0..5 (K layout: the way back to the identity layout)"""
        return self.first_fit(remaining, self.full & ~LOW, LOW)[1]


def schedule(masks, n_local, leave_ratio=0.8):
    """-> list of (logical qubit mask, [gate indices in program order], transposing).  The layout is the identity
    before the first and after the last block."""
    s = _Sched(masks, n_local)
    remaining = list(range(len(masks)))
    out = []

    def drop(rem, tk):
        t = set(tk)
        return [i for i in rem if i not in t]

    idle = 0
    while remaining:
        if idle >= 3:
            return None   # a gate needs more free physical bits than this state has beside the parked ones
        if s.K is None:
            ch, tk = s.best_block(remaining, LOW)
            needs_low = any(masks[i] & LOW for i in remaining)
            if not tk:
                # every frontier gate touches a parked qubit: park the six qubits used furthest in the future instead
                nxt = {}
                for pos, i in enumerate(remaining):
                    for b in range(K_BLOCK, n_local):
                        if (masks[i] >> b) & 1 and b not in nxt:
                            nxt[b] = pos
                order = sorted(range(K_BLOCK, n_local), key=lambda b: -nxt.get(b, 1 << 60))
                ch = 0
                for b in order:
                    if bin(ch).count("1") < K_BLOCK and pad_kpos([x for x in range(n_local) if (ch >> x) & 1] + [b], n_local):
                        ch |= 1 << b
            kpos = pad_kpos([b for b in range(n_local) if (ch >> b) & 1], n_local)
            assert kpos is not None
            if needs_low:
                out.append((kpos, tk, True))
                s.K, s.parked = kpos, sum(1 << b for b in kpos)
                s._feas.clear()
            else:
                out.append((kpos, tk, False))
            remaining = drop(remaining, tk)
            idle = 0 if tk else idle + 1
        else:
            chA, tkA = s.best_block(remaining, s.parked)
            tkB = s.exact_low(remaining)
            remB = drop(remaining, tkB)
            # what the identity layout would offer next
            saved = (s.K, s.parked)
            s.K, s.parked = None, LOW
            s._feas.clear()
            _, tkB2 = s.best_block(remB, LOW)
            s.K, s.parked = saved
            s._feas.clear()
            gA, gB = len(tkA), (len(tkB) + len(tkB2)) / 2
            leave = gA == 0 or gB >= gA * leave_ratio
            if leave:
                out.append((list(s.K), tkB, True))
                remaining = remB
                idle = 0 if tkB else idle + 1
                s.K, s.parked = None, LOW
                s._feas.clear()
            else:
                kpos = pad_kpos([s.phys(b) for b in range(n_local) if (chA >> b) & 1], n_local)
                out.append((kpos, tkA, False))
                remaining = drop(remaining, tkA)
                idle = 0
    if s.K is not None:
        out.append((list(s.K), [], True))
    return _with_layouts(out)


def _with_layouts(blocks):
    """annotate every block with the logical -> physical map in force while it runs: list of (kpos, gates, transposing, K)"""
    res, K = [], None
    for kpos, tk, tr in blocks:
        res.append((kpos, tk, tr, None if K is None else list(K)))
        if tr:
            K = list(kpos) if K is None else None
    assert K is None
    return res


def _next_use(masks, remaining, n_local):
    nxt = {}
    for at, i in enumerate(remaining):
        m = masks[i]
        b = 0
        while m >> b:
            if (m >> b) & 1 and b not in nxt:
                nxt[b] = at
            b += 1
        if len(nxt) == n_local:
            break
    return nxt


def schedule_rotating(masks, n_local):
    """Rotating write-back: -> list of (kpos in pattern order, [gate indices], transposing, where) with `where` = logical
    qubit -> block-local bit; the last entries are the gate-free passes that bring every qubit home.  Any block may
    transpose, which parks ITS six qubits in index bits 0-5 and frees the six parked before; it does when its qubits
    (after the block) stay idle longer than the parked ones."""
    if n_local < 3 * K_BLOCK:
        return None      # the way home needs six un-parked qubits outside 0-5 (smaller states: in-place schedule)
    s = _Sched(masks, n_local, rotating=True)
    remaining = list(range(len(masks)))
    out = []   # (logical qubit mask, gates, transposing, fixed pattern order or None)

    def drop(rem, tk):
        t = set(tk)
        return [i for i in rem if i not in t]

    def fill(ch, rem, avoid):
        """pad the chosen set to six qubits with un-parked qubits, those used furthest in the future first"""
        nxt = _next_use(masks, rem, n_local)
        order = sorted((b for b in range(n_local) if not ((ch | avoid) >> b) & 1), key=lambda b: -nxt.get(b, 1 << 60))
        for b in order:
            if bin(ch).count("1") >= K_BLOCK:
                break
            ch |= 1 << b
        return ch

    while remaining:
        ch, tk = s.best_block(remaining, s.parked)
        if not tk:
            # every frontier gate touches a parked qubit: a gate-free exchange that parks the idlest six instead
            ch = fill(0, remaining, s.parked)
            out.append((ch, [], True, None))
            s.parked = ch
            continue
        rem2 = drop(remaining, tk)
        chB = fill(ch, rem2, s.parked)
        nxt = _next_use(masks, rem2, n_local)
        idle = lambda m: sum(min(nxt.get(b, len(rem2)), len(rem2)) for b in range(n_local) if (m >> b) & 1)
        stay, swap = idle(s.parked), idle(chB)   # park whichever six qubits stay idle longer (summed distance to next use)
        if swap > stay:
            out.append((chB, tk, True, None))
            s.parked = chB
        else:
            out.append((ch, tk, False, None))
        remaining = rem2
    # the way home for the parked set: logical 0-5 into index bits 0-5, in order
    # (a transposing block on exactly the qubits 0-5 always uses the pattern order 0..5: _rotating_layouts)
    if s.parked & LOW and s.parked != LOW:
        ch = fill(0, [], s.parked | LOW)
        out.append((ch, [], True, None))
        s.parked = ch
    if s.parked != LOW:
        out.append((LOW, [], True, None))
    return _rotating_layouts(out, n_local)


def _rotate(pos, qs, transposing, n_local):
    """logical -> physical map after a rotating block on the logical qubits qs (pattern order)"""
    inv = {p: l for l, p in enumerate(pos)}
    new = [None] * n_local
    for i, q in enumerate(qs):
        new[q] = i if transposing else K_BLOCK + i
    for b in range(K_BLOCK):
        new[inv[b]] = K_BLOCK + b if transposing else b
    r = 2 * K_BLOCK
    inblock = set(qs)
    for p in range(K_BLOCK, n_local):
        if inv[p] not in inblock:
            new[inv[p]] = r
            r += 1
    return new


def _rotating_layouts(blocks, n_local):
    pos = list(range(n_local))
    res = []

    def emit(qs, tk, tr):
        nonlocal pos
        res.append(([pos[q] for q in qs], tk, tr, {q: i for i, q in enumerate(qs)}))
        pos = _rotate(pos, qs, tr, n_local)

    for ch, tk, tr, order in blocks:
        if order is None and tr and ch == LOW:
            order = list(range(K_BLOCK))     # the parked qubits 0-5 come home in order
        if order is None:
            qs = [b for b in range(n_local) if (ch >> b) & 1]
            # pad with un-parked neighbours (physically adjacent first: fewer TMA pieces), then anything un-parked
            while len(qs) < K_BLOCK:
                have = {pos[q] for q in qs}
                inv = {p: l for l, p in enumerate(pos)}
                cands = [p for p in range(K_BLOCK, n_local) if p not in have]
                cands.sort(key=lambda p: (0 if (p - 1 in have or p + 1 in have) else 1, p))
                qs.append(inv[cands[0]])
            qs.sort(key=lambda q: pos[q])
        else:
            qs = list(order)
        emit(qs, tk, tr)
    assert all(pos[b] == b for b in range(K_BLOCK))
    # sort the qubits >= 6 home: every pass moves six of them to the front, the top chunk first; the qubits T.. that are
    # already in ascending relative order are never moved (they end up on top, in order)
    T = n_local
    while T - 1 > K_BLOCK - 1 and (T == n_local or pos[T - 1] < pos[T]):
        T -= 1
    if any(pos[l] != l for l in range(n_local)):
        todo = list(range(K_BLOCK, T))
        chunks = []
        while todo:
            chunks.append(todo[:K_BLOCK])
            todo = todo[K_BLOCK:]
        nxt, low = T, K_BLOCK
        while len(chunks[-1]) < K_BLOCK:      # pad the top chunk with the lowest qubits of the sorted rest, or (no rest
            if nxt < n_local:                 # left) with qubits that a later pass moves again anyway
                chunks[-1].append(nxt)
                nxt += 1
            else:
                chunks[-1].append(low)
                low += 1
        for chunk in reversed(chunks):
            emit(chunk, [], False)
    assert pos == list(range(n_local)), pos
    return res


class BlockStage:
    """A run of gates executed as dense 6-qubit blocks (no autograd)."""
    kind = "blocks"

    def __init__(self):
        self.blocks = None    # int32 [n_blocks, KB_WORDS]
        self.ops = None       # int32 [n_ops, KO_WORDS]
        self.consts = None    # float32
        self.sh_recs = []     # queue indices of the shared-parameter records, parameter order
        self.P_sh = 0
        self.n_ops = self.n_blocks = 0
        self.n_rotating = 0   # blocks with TQB_KBF_ROTATE: each moves the state to the other buffer


def _full_const(op):
    """constant matrix of a lowered op in the block builder's full complex form (float32 pairs)"""
    if op.type == C.T_HS:
        return _H.view(np.float32).reshape(-1)
    if op.type == C.T_REAL:
        return np.asarray(op.const, dtype=np.float32).astype(np.complex64).view(np.float32).reshape(-1)
    return np.asarray(op.const, dtype=np.float32).reshape(-1)


# Relative cost of one in-place block pass by placement (profiles/r02_block6_kpos_sweep.txt, n = 28: 0.80-0.85 ms when
# a block bit is <= 11 or at most two are >= 19, up to 1.37 ms when all six are high); a rotating pass always writes
# contiguous tiles (ROTATING_COST, measured: profiles/r02_block6_rotate_sweep.txt).
ROTATING_COST = 1.0
MODE = os.environ.get("TQB200_BLOCK_WRITEBACK", "auto")   # auto | inplace | rotate


def inplace_cost(kpos):
    if min(kpos) <= 11:
        return 1.0
    high = sum(1 for b in kpos if b >= 19)
    return 1.04 if high <= 2 else (1.2 if high <= 4 else 1.5)


def plan_blocks(masks, n_local):
    """-> (list of (kpos, gates, flags, where), estimated cost in units of one fast pass) for the cheaper of the two
    write-back disciplines, or None"""
    cands = []
    if MODE != "rotate":
        sched = schedule(masks, n_local)
        if sched:
            ent = []
            for kpos, gates, tr, K in sched:
                at = {p: i for i, p in enumerate(kpos)}
                where = {l: at[p] for l in range(n_local)
                         for p in [l if K is None else (K[l] if l < K_BLOCK else l)] if p in at}
                ent.append((kpos, gates, C.KBF_TRANSPOSE if tr else 0, where))
            cands.append((sum(inplace_cost(e[0]) for e in ent), ent))
    if MODE != "inplace":
        sched = schedule_rotating(masks, n_local)
        if sched:
            ent = [(kpos, gates, C.KBF_ROTATE | (C.KBF_TRANSPOSE if tr else 0), where) for kpos, gates, tr, where in sched]
            cands.append((ROTATING_COST * len(ent), ent))
    if not cands:
        return None
    cost, ent = min(cands, key=lambda c: c[0])
    return ent, cost


def build_block_stage(lops, n_local, min_avg=0.0):
    """lops: fusable lowered ops (plan.LOp) in program order -> BlockStage, or None when the schedule averages fewer
    than `min_avg` gates per block (the register-tile pass kernel is the better tool then)."""
    lmask = (1 << n_local) - 1
    masks = [op.qmask & lmask for op in lops]
    planned = plan_blocks(masks, n_local)
    if planned is None or len(lops) / planned[1] < min_avg:
        return None
    sched = planned[0]
    st = BlockStage()
    st.n_rotating = sum(1 for e in sched if e[2] & C.KBF_ROTATE)
    for op in lops:
        if op.pmode == "sh":
            op.poff = st.P_sh
            st.P_sh += op.nparam
            st.sh_recs.append(op.rec)
    blocks, words, consts, coff = [], [], [], 0
    for kpos, gates, flags, where in sched:

        def bl(l):
            return where[l]

        hdr = [0] * C.KB_WORDS
        hdr[C.KB_OPFIRST], hdr[C.KB_NOPS], hdr[C.KB_FLAGS] = len(words), len(gates), flags
        for i, p in enumerate(kpos):
            hdr[C.KB_KPOS + i] = p
        blocks.append(hdr)
        for gi in gates:
            op = lops[gi]
            w = [0] * C.KO_WORDS
            cm = cv = pm = rcm = rcv = rpm = 0
            b = 0
            while (op.cmask >> b) or (op.pmask >> b):
                if b >= n_local:
                    rb = b - n_local
                    rcm |= ((op.cmask >> b) & 1) << rb
                    rcv |= ((op.cval >> b) & 1) << rb
                    rpm |= ((op.pmask >> b) & 1) << rb
                else:
                    if (op.cmask >> b) & 1:
                        cm |= 1 << bl(b)
                        cv |= ((op.cval >> b) & 1) << bl(b)
                    if (op.pmask >> b) & 1:
                        pm |= 1 << bl(b)
                b += 1
            if op.type == C.T_PERM:
                form, t0, t1 = C.KF_X, bl(op.targets[0]), 0
            elif op.type == C.T_DIAG:
                form, t0, t1 = C.KF_DIAG, 0, 0
            elif op.type == C.T_U2Q:
                form, t0, t1 = C.KF_U2, bl(op.targets[1]), bl(op.targets[0])
            else:
                form, t0, t1 = C.KF_U1, bl(op.targets[0]), 0
            w[C.KO_BITS] = t0 | (t1 << 8)
            w[C.KO_MASKS] = cm | (cv << 8) | (pm << 16)
            w[C.KO_RCM], w[C.KO_RCV], w[C.KO_RPM] = rcm, rcv, rpm
            if op.kind == C.K_CONST:
                w[C.KO_W0] = form
                w[C.KO_KIND] = C.K_CONST
                if form != C.KF_X:
                    c = _full_const(op)
                    assert len(c) == {C.KF_U1: 8, C.KF_DIAG: 4, C.KF_U2: 32}[form]
                    w[C.KO_OFF] = coff
                    consts.append(c)
                    coff += len(c)
            else:
                w[C.KO_W0] = form | (op.nparam << 4) | ((1 << 8) if op.inverse else 0)
                w[C.KO_KIND], w[C.KO_OFF] = op.kind, op.poff
            words.append(w)
    st.blocks = np.array(blocks, dtype=np.int32).reshape(-1, C.KB_WORDS)
    st.ops = np.array(words if words else [[0] * C.KO_WORDS], dtype=np.int32).reshape(-1, C.KO_WORDS)
    st.consts = np.concatenate(consts + [np.zeros(4, np.float32)]).astype(np.float32)
    st.n_ops, st.n_blocks = len(lops), len(blocks)
    return st
