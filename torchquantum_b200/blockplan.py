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


QUAD_MIN = 8       # fewest commuting diagonal gates worth a quadratic-phase pass of their own
QUAD_RATIO = 1.5   # ... and they must outnumber what the best block would absorb by this factor


def schedule_rotating(masks, n_local, diag=None, swaps=None, final=None):
    """Rotating write-back: -> list of entries in execution order,
         ("block", kpos in pattern order, [gate indices], transposing, where, perm)   where = logical qubit -> block-local
                                            bit while the gates act, perm = swaps of block-local bits applied after them
         ("quad", [gate indices], pos)      commuting 1- / 2-qubit diagonal gates as ONE quadratic-phase pass; pos = logical
                                            -> physical bit while it runs
       the last blocks are the gate-free passes that bring every qubit home.  Any block may transpose, which parks ITS
       six qubits in index bits 0-5 and frees the six parked before; it does when its qubits (after the block) stay
       idle longer than the parked ones.
       diag[i]: gate i is a diagonal of degree <= 2 (quad_terms); such gates commute with each other, and when the
       dependency frontier holds many more of them than the best block would absorb they run as a quadratic-phase pass
       (the twiddle stages of a QFT).  swaps[i] = (a, b): gate i is a plain SWAP; at the frontier it only relabels the two
       logical qubits (the layout is sorted at the end anyway).
       final: a list -> the stage does NOT bring the qubits home; the logical -> physical map it ends with is appended
       (callers that track the layout themselves: the sharded device)."""
    if n_local < 3 * K_BLOCK:
        return None      # the way home needs six un-parked qubits outside 0-5 (smaller states: in-place schedule)
    s = _Sched(masks, n_local, rotating=True)
    remaining = list(range(len(masks)))
    out = []   # ("block", logical qubit mask, gates, transposing) | ("quad", gates) | ("relabel", a, b)
    full = (1 << n_local) - 1

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

    def frontier(rem, want):
        """gates of `rem` flagged by want(i) that no earlier un-flagged gate blocks (flagged gates commute with each other)"""
        blocked, got = 0, []
        for i in rem:
            if want(i) and not (masks[i] & blocked):
                got.append(i)
            elif not want(i):
                blocked |= masks[i]
                if blocked == full:
                    break
        return got

    while remaining:
        if swaps is not None:
            # a SWAP at the frontier of everything (diagonals included) is a relabelling
            blocked, done = 0, []
            for i in remaining:
                if swaps[i] is not None and not (masks[i] & blocked):
                    a, b = swaps[i]
                    out.append(("relabel", a, b))
                    pa, pb = (s.parked >> a) & 1, (s.parked >> b) & 1
                    if pa != pb:
                        s.parked ^= (1 << a) | (1 << b)
                    done.append(i)
                else:
                    blocked |= masks[i]
                    if blocked == full:
                        break
            if done:
                remaining = drop(remaining, done)
                if not remaining:
                    break
        ch, tk = s.best_block(remaining, s.parked)
        if diag is not None:
            D = frontier(remaining, lambda i: diag[i])
            if len(D) >= QUAD_MIN and len(D) > QUAD_RATIO * len(tk):
                out.append(("quad", D))
                remaining = drop(remaining, D)
                continue
        if not tk:
            # every frontier gate touches a parked qubit: a gate-free exchange that parks the idlest six instead
            ch = fill(0, remaining, s.parked)
            out.append(("block", ch, [], True))
            s.parked = ch
            continue
        rem2 = drop(remaining, tk)
        chB = fill(ch, rem2, s.parked)
        nxt = _next_use(masks, rem2, n_local)
        idle = lambda m: sum(min(nxt.get(b, len(rem2)), len(rem2)) for b in range(n_local) if (m >> b) & 1)
        stay, swap = idle(s.parked), idle(chB)   # park whichever six qubits stay idle longer (summed distance to next use)
        if swap > stay:
            out.append(("block", chB, tk, True))
            s.parked = chB
        else:
            out.append(("block", ch, tk, False))
        remaining = rem2
    return _rotating_layouts(out, n_local, final)


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


def _rotating_layouts(entries, n_local, final=None):
    pos = list(range(n_local))
    res = []

    def emit(qs, tk, tr):
        """qs: the block's qubits in the order they are to ARRIVE (pattern bit i of the written tile = qs[i]).  The tile is
        always gathered in ascending physical order (longest TMA runs); a different arrival order is a permutation of the
        block-local bits multiplied into the block unitary (0 / 1 entries: exact), given as swaps after the gates."""
        nonlocal pos
        load = sorted(qs, key=lambda q: pos[q])
        at = {q: i for i, q in enumerate(load)}       # where each qubit sits while the gates act
        cur, perm = list(load), []
        for i, q in enumerate(qs):                     # selection sort of the block-local bits into arrival order
            j = cur.index(q)
            if j != i:
                perm.append((i, j))
                cur[i], cur[j] = cur[j], cur[i]
        res.append(("block", [pos[q] for q in load], tk, tr, at, perm))
        pos = _rotate(pos, qs, tr, n_local)

    for ent in entries:
        if ent[0] == "relabel":
            _, a, b = ent
            pos[a], pos[b] = pos[b], pos[a]
            continue
        if ent[0] == "quad":
            res.append(("quad", ent[1], list(pos)))
            continue
        _, ch, tk, tr = ent
        qs = [b for b in range(n_local) if (ch >> b) & 1]
        # pad with un-parked neighbours (physically adjacent first: fewer TMA pieces), then anything un-parked
        while len(qs) < K_BLOCK:
            have = {pos[q] for q in qs}
            inv = {p: l for l, p in enumerate(pos)}
            cands = [p for p in range(K_BLOCK, n_local) if p not in have]
            cands.sort(key=lambda p: (0 if (p - 1 in have or p + 1 in have) else 1, p))
            qs.append(inv[cands[0]])
        qs.sort(key=lambda q: pos[q])
        emit(qs, tk, tr)
    if final is not None:
        final.append(list(pos))
        return res
    # the way home for the parked set: logical 0-5 into index bits 0-5, in order
    if pos[:K_BLOCK] != list(range(K_BLOCK)):
        if any(pos[l] < K_BLOCK for l in range(K_BLOCK)):
            # some of the qubits 0-5 are parked (but not all of them in place): park six others first (the highest ones)
            emit(sorted((l for l in range(K_BLOCK, n_local) if pos[l] >= K_BLOCK), key=lambda l: pos[l])[-K_BLOCK:], [], True)
        emit(list(range(K_BLOCK)), [], True)
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
        self.items = []       # execution order: ("blocks", first block, count) | ("quad", index into quads)
        self.quads = []       # QuadRecipe of every quadratic-phase pass
        self.final_pos = None # logical -> physical bit after the stage when it does not bring the qubits home


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


QUAD_COST = 0.95     # a quadratic-phase pass streams the state once (no tensor-core work)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex64).view(np.float32).reshape(-1)


def quad_terms(op, n_local):
    """A lowered diagonal op as terms of a quadratic form over index bits, or None when it is not one of degree <= 2.
    The op multiplies by d[parity(x & pmask)] where (x & cmask) == cval, with d_k = exp(i theta_k).  -> list of
    (bits, w0, w1): the coefficient of prod_{b in bits} x_b grows by w0 * theta_0 + w1 * theta_1.  Bits are the op's own
    (logical local bits, rank bits at >= n_local)."""
    if op.type != C.T_DIAG or op.pmode not in (None, "sh") or op.kind == C.K_EXTERN or not op.unitary:
        return None
    if op.postx is not None or op.nounapply or op.cval != op.cmask:
        return None
    if op.kind == C.K_CONST:
        d = np.asarray(op.const, dtype=np.float64)
        if abs(d[0] ** 2 + d[1] ** 2 - 1) > 1e-5 or abs(d[2] ** 2 + d[3] ** 2 - 1) > 1e-5:
            return None
    elif op.kind not in (C.K_RZ, C.K_PHASE, C.K_GPHASE):
        return None
    bits = lambda m: tuple(b for b in range(m.bit_length()) if (m >> b) & 1)
    ctrl, free = bits(op.cmask), bits(op.pmask & ~op.cmask)
    if len(ctrl) + len(free) > 2:
        return None
    flip = bin(op.pmask & op.cmask).count("1") & 1        # parity bits that are also controls are set on the support
    a0 = (0.0, 1.0) if flip else (1.0, 0.0)              # factor on the support where the free parity is even ...
    dl = (1.0, -1.0) if flip else (-1.0, 1.0)            # ... and the step to where it is odd, as weights of (theta_0, theta_1)
    terms = [(ctrl, a0[0], a0[1])]
    if len(free) == 1:
        terms.append((tuple(sorted(ctrl + free)), dl[0], dl[1]))
    elif len(free) == 2:      # x_a xor x_b = x_a + x_b - 2 x_a x_b
        terms += [((free[0],), dl[0], dl[1]), ((free[1],), dl[0], dl[1]), (free, -2 * dl[0], -2 * dl[1])]
    return terms


def _thetas(op):
    """(theta_0, theta_1) of a diagonal op: constants (radians), or the weights of its one parameter"""
    sgn = -1.0 if op.inverse else 1.0
    if op.kind == C.K_CONST:    # (inverse already folded into the constant)
        d = np.asarray(op.const, dtype=np.float64)
        return ("const", float(np.arctan2(d[1], d[0])), float(np.arctan2(d[3], d[2])))
    if op.kind == C.K_RZ:
        return ("param", -0.5 * sgn, 0.5 * sgn)
    if op.kind == C.K_PHASE:
        return ("param", 0.0, sgn)
    return ("param", sgn, sgn)      # K_GPHASE


class QuadRecipe:
    """coefficients (in turns) of one quadratic-phase pass: base + sum_k weight_k * params_sh[pidx_k] at slot idx_k"""

    def __init__(self, lops, gates, pos, n_local):
        base = np.zeros(C.QP_WORDS, dtype=np.float64)
        idx, w, pidx = [], [], []
        top = n_local
        for gi in gates:
            op = lops[gi]
            kind, t0, t1 = _thetas(op)
            for bits, w0, w1 in quad_terms(op, n_local):
                phys = sorted(pos[b] if b < n_local else b for b in bits)
                top = max([top] + [b + 1 for b in phys])
                slot = 0 if not phys else (1 + phys[0] if len(phys) == 1 else 1 + C.QP_BITS + phys[0] * C.QP_BITS + phys[1])
                if kind == "const":
                    base[slot] += (w0 * t0 + w1 * t1) / (2 * np.pi)
                else:
                    idx.append(slot)
                    w.append((w0 * t0 + w1 * t1) / (2 * np.pi))
                    pidx.append(op.poff)
        assert top <= C.QP_BITS
        self.n_bits, self.n_gates = top, len(gates)
        self.base = base
        self.idx = np.array(idx, dtype=np.int64)
        self.w = np.array(w, dtype=np.float64)
        self.pidx = np.array(pidx, dtype=np.int64)


def plan_blocks(masks, n_local, diag=None, swaps=None, come_home=True):
    """-> (entries, estimated cost in units of one fast pass, final layout or None = identity) for the cheaper of the two
    write-back disciplines, or None.  come_home=False: a rotating schedule may leave the qubits where they are.
    entries: ("block", kpos, gates, flags, where, perm) | ("quad", gates, pos), see schedule_rotating; perm: swaps of
    block-local bits multiplied into the block unitary after the gates"""
    cands = []
    if MODE != "rotate":
        sched = schedule(masks, n_local)
        if sched:
            ent = []
            for kpos, gates, tr, K in sched:
                at = {p: i for i, p in enumerate(kpos)}
                where = {l: at[p] for l in range(n_local)
                         for p in [l if K is None else (K[l] if l < K_BLOCK else l)] if p in at}
                ent.append(("block", kpos, gates, C.KBF_TRANSPOSE if tr else 0, where, []))
            cands.append((sum(inplace_cost(e[1]) for e in ent), ent, None))
    if MODE != "inplace":
        final = None if come_home else []
        sched = schedule_rotating(masks, n_local, diag, swaps, final)
        if sched:
            ent = [("block", e[1], e[2], C.KBF_ROTATE | (C.KBF_TRANSPOSE if e[3] else 0), e[4], e[5]) if e[0] == "block" else e
                   for e in sched]
            cands.append((sum(ROTATING_COST if e[0] == "block" else QUAD_COST for e in ent), ent, final[0] if final else None))
    if not cands:
        return None
    cost, ent, fin = min(cands, key=lambda c: c[0])
    if fin is not None and fin == list(range(n_local)):
        fin = None
    return ent, cost, fin


def build_block_stage(lops, n_local, min_avg=0.0, come_home=True):
    """lops: fusable lowered ops (plan.LOp) in program order -> BlockStage, or None when the schedule averages fewer
    than `min_avg` gates per pass (the register-tile pass kernel is the better tool then).  come_home=False: the stage
    may end in another qubit layout (BlockStage.final_pos: logical -> physical bit)."""
    lmask = (1 << n_local) - 1
    masks = [op.qmask & lmask for op in lops]
    diag = [quad_terms(op, n_local) is not None for op in lops]
    swaps = [(op.targets[0], op.targets[1]) if (op.type == C.T_U2Q and op.kind == C.K_CONST and not op.cmask
                                                 and np.array_equal(np.asarray(op.const, dtype=np.float32).reshape(-1), _SWAP))
             else None for op in lops]
    planned = plan_blocks(masks, n_local, diag if any(diag) else None, swaps if any(sw is not None for sw in swaps) else None,
                          come_home)
    if planned is None or len(lops) / planned[1] < min_avg:
        return None
    sched = planned[0]
    st = BlockStage()
    st.final_pos = planned[2]
    for op in lops:
        if op.pmode == "sh":
            op.poff = st.P_sh
            st.P_sh += op.nparam
            st.sh_recs.append(op.rec)
    blocks, words, consts, coff, swap_off = [], [], [], 0, None
    for ent in sched:
        if ent[0] == "quad":
            st.items.append(("quad", len(st.quads)))
            st.quads.append(QuadRecipe(lops, ent[1], ent[2], n_local))
            continue
        _, kpos, gates, flags, where, perm = ent
        if st.items and st.items[-1][0] == "blocks":
            st.items[-1] = ("blocks", st.items[-1][1], st.items[-1][2] + 1)
        else:
            st.items.append(("blocks", len(blocks), 1))

        def bl(l):
            return where[l]

        hdr = [0] * C.KB_WORDS
        hdr[C.KB_OPFIRST], hdr[C.KB_NOPS], hdr[C.KB_FLAGS] = len(words), len(gates) + len(perm), flags
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
        for i, j in perm:      # arrival order of the block's qubits: constant SWAPs of block-local bits
            if swap_off is None:
                swap_off = coff
                consts.append(_SWAP)
                coff += len(_SWAP)
            w = [0] * C.KO_WORDS
            w[C.KO_BITS] = i | (j << 8)
            w[C.KO_W0], w[C.KO_KIND], w[C.KO_OFF] = C.KF_U2, C.K_CONST, swap_off
            words.append(w)
    st.blocks = np.array(blocks, dtype=np.int32).reshape(-1, C.KB_WORDS)
    st.ops = np.array(words if words else [[0] * C.KO_WORDS], dtype=np.int32).reshape(-1, C.KO_WORDS)
    st.consts = np.concatenate(consts + [np.zeros(4, np.float32)]).astype(np.float32)
    st.n_ops, st.n_blocks = len(lops), len(blocks)
    st.n_rotating = int(sum(1 for h in blocks if h[C.KB_FLAGS] & C.KBF_ROTATE))
    return st
