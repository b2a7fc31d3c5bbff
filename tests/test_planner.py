"""Host planner unit tests (no kernels): lowering choices, folds and program structure of torchquantum_b200/plan.py."""
import numpy as np

from torchquantum_b200 import _codes as C
from torchquantum_b200 import plan as P


def _types(plan):
    out = []
    for st in plan.stages:
        for pp in st.passes:
            w = pp.prog
            for si in range(w[C.H_NSEG]):
                off = w[C.H_SEGOFF + si]
                base = off + C.SEG_HDR_WORDS + w[off + C.S_NGROUPS] * C.GROUP_WORDS
                for k in range(w[off + C.S_NOPS]):
                    w0 = int(w[base + k * C.OP_WORDS]) & 0xFFFFFFFF
                    out.append(((w0 >> 2) & 15, bool(w0 & C.F_POSTX)))
    return out


def _key(ops):
    return tuple((name, tuple(wires), False, pm, grad, 0) for name, wires, pm, grad in ops)


def test_grad_programs_use_plain_forms_and_forward_only_programs_scaled_ones():
    ops = [("rx", [0], "sh", True), ("ry", [1], "sh", True), ("rz", [2], "sh", True), ("hadamard", [3], None, False),
           ("rx", [1], "sh", True), ("ry", [2], "sh", True)]
    grad = {t for t, _ in _types(P.build_plan(4, _key(ops), True))}
    assert grad == {C.T_RX, C.T_REAL, C.T_DIAG}
    fwd_ops = [(n, w, pm, False) for n, w, pm, _ in ops]
    fwd = {t for t, _ in _types(P.build_plan(4, _key(fwd_ops), False))}
    # forward-only: rotations keep their plain (normalised) forms - the ratio forms overflowed deep passes and drifted
    # the norm (plan.SCALED_ROT) - and a lone H is not worth a deferred factor (SCALED_MIN)
    assert fwd == {C.T_RX, C.T_REAL, C.T_DIAG}
    many_h = fwd_ops + [("hadamard", [w], None, False) for w in (0, 1, 2)]
    assert C.T_HS in {t for t, _ in _types(P.build_plan(4, _key(many_h), False))}
    # more scaled ops in one pass than SCALED_MAX: plain H again (un-normalised intermediates must stay in range)
    deep = [("hadamard", [k % 4], None, False) for k in range(P.SCALED_MAX + 4)]
    assert {t for t, _ in _types(P.build_plan(4, _key(deep), False))} == {C.T_REAL}


def test_cnot_sandwich_folds_into_one_parity_diagonal():
    ops = [("cnot", [0, 1], None, False), ("rz", [1], "sh", True), ("cnot", [0, 1], None, False)]
    plan = P.build_plan(3, _key(ops), True)
    assert [t for t, _ in _types(plan)] == [C.T_DIAG]


def test_cnot_folds_into_the_preceding_gate_in_forward_only_programs():
    ops = [("rx", [0], "sh", False), ("ry", [1], "sh", False), ("rz", [2], "sh", False), ("cnot", [0, 1], None, False),
           ("ry", [2], "sh", False)]
    types = _types(P.build_plan(3, _key(ops), False))
    assert (C.T_REAL, True) in types, types   # ry absorbs the CNOT on its target
    # with autograd the CNOT stays a register swap of its own
    gtypes = _types(P.build_plan(3, _key([(n, w, pm, pm is not None) for n, w, pm, _ in ops]), True))
    assert (C.T_PERM, False) in gtypes


def test_qft_rungs_fold_into_ladders():
    n = 9
    ops = []
    for top in range(n):
        ops.append(("hadamard", [top], None, False))
        for w in range(top + 1, n):
            ops.append(("cu1", [w, top], "sh", False))
    plan = P.build_plan(n, _key(ops), False)
    types = [t for t, _ in _types(plan)]
    assert types.count(C.T_LADDER) == n - 1 - (P.LADDER_MIN - 1)   # the last LADDER_MIN - 1 rung sets are too short
    assert plan.stages[0].P_sh == n * (n - 1) // 2                  # every rung keeps its parameter slot
    # programs that run backward keep one masked diagonal per rung
    gplan = P.build_plan(n, _key([(nm, w, pm, pm is not None) for nm, w, pm, _ in ops]), True)
    assert all(t != C.T_LADDER for t, _ in _types(gplan))


def test_streamed_states_are_cut_into_passes_of_at_most_seven_free_bits():
    n = 20
    ops = [("rx", [w], "sh", False) for w in range(n)]
    plan = P.build_plan(n, _key(ops), False)
    st = plan.stages[0]
    assert len(st.passes) == 3 and all(pp.T == P.T_MULTI for pp in st.passes)   # 15 high bits / 7 per pass
    for pp in st.passes:
        assert pp.tile_bits[:P.L_MIN] == list(range(P.L_MIN))                    # 256-byte runs
        assert int(pp.prog[C.H_WORDS]) == len(pp.prog)
    covered = set()
    for pp in st.passes:
        covered |= set(pp.tile_bits)
    assert covered == set(range(n))


def test_segments_respect_staging_capacity():
    n = 6
    rng = np.random.default_rng(0)
    ops = [("u3", [int(rng.integers(n))], "sh", True) for _ in range(200)]
    plan = P.build_plan(n, _key(ops), True)
    for pp in plan.stages[0].passes:
        w = pp.prog
        assert w[C.H_G_MAX] <= C.SEG_MAX_G and w[C.H_SH_MAX] <= C.SEG_MAX_SH
        for si in range(w[C.H_NSEG]):
            off = w[C.H_SEGOFF + si]
            assert w[off + C.S_NOPS] <= C.SEG_MAX_OPS and w[off + C.S_G_LEN] <= C.SEG_MAX_G


# ---- block scheduler (blockplan.py): in-place and rotating write-back ----

def _brick_masks(n, depth):
    masks = []
    for d in range(depth):
        masks += [1 << q for q in range(n)]
        masks += [3 << q for q in range(d % 2, n - 1, 2)]
    return masks


def _random_masks(n, count, seed):
    rng = np.random.default_rng(seed)
    masks = []
    for _ in range(count):
        k = int(rng.integers(1, 4))
        qs = rng.choice(n, size=k, replace=False)
        masks.append(int(sum(1 << int(q) for q in qs)))
    return masks


def _check_rotating_schedule(masks, n, diag=None, swaps=None):
    """Replays a rotating schedule against the kernel's contract (tqb_apply_blocks): every gate exactly once, inside its
    block, after everything it depends on (diagonal gates may pass each other, a relabelled SWAP exchanges the two
    wires of all later gates); the layout the scheduler tracks matches the replay and ends as the identity."""
    from torchquantum_b200 import blockplan as B

    sched = B.schedule_rotating(masks, n, diag, swaps)
    assert sched is not None
    pos = list(range(n))            # DATA item (initial wire) d -> physical bit
    wire = list(range(n))           # data item currently on logical wire w (SWAPs exchange them)
    done = set()
    is_diag = lambda g: diag is not None and diag[g]
    is_swap = lambda g: swaps is not None and swaps[g] is not None

    def retire_swaps():
        # a SWAP whose predecessors are all done has been turned into a relabelling by the scheduler
        progress = True
        while progress:
            progress = False
            for g in range(len(masks)):
                if g not in done and is_swap(g) and all(h in done for h in range(g) if masks[h] & masks[g]):
                    a, b = swaps[g]
                    wire[a], wire[b] = wire[b], wire[a]
                    done.add(g)
                    progress = True

    def check_deps(g):
        for h in range(g):
            if masks[h] & masks[g] and h not in done:
                assert is_diag(h) and is_diag(g), (h, g)     # only diagonal gates may overtake each other

    n_quads = 0
    for ent in sched:
        retire_swaps()
        if ent[0] == "quad":
            _, gates, qpos = ent
            assert qpos == [pos[wire[w]] for w in range(n)]
            for g in gates:
                assert is_diag(g)
                check_deps(g)
            done.update(gates)
            n_quads += 1
            continue
        _, kpos, gates, tr, where, perm = ent
        assert len(set(kpos)) == 6 and all(6 <= p < n for p in kpos)
        assert kpos == sorted(kpos)                                   # tiles are gathered in ascending physical order
        inv = {pos[wire[w]]: w for w in range(n)}          # physical bit -> logical wire
        qs = [inv[p] for p in kpos]
        assert all(where[q] == i for i, q in enumerate(qs))          # block-local bit i <-> physical kpos[i]
        blockmask = sum(1 << q for q in qs)
        for g in gates:
            assert not is_swap(g) or True
            assert masks[g] & ~blockmask == 0                          # every gate of the block acts inside it
            check_deps(g)
            done.add(g)
        # layout after the pass: block bits -> 6..11 (0..5 when transposing, the parked six then at 6..11), rest above
        new = {}
        pinv = {p: d for d, p in enumerate(pos)}
        arrive = list(kpos)                                           # block-local bit i holds the data of physical arrive[i]
        for i, j in perm:
            arrive[i], arrive[j] = arrive[j], arrive[i]
        for i, p in enumerate(arrive):
            new[pinv[p]] = i if tr else 6 + i
        for b in range(6):
            new[pinv[b]] = 6 + b if tr else b
        r = 12
        for p in range(6, n):
            if p not in kpos:
                new[pinv[p]] = r
                r += 1
        pos = [new[d] for d in range(n)]
    retire_swaps()
    assert done == set(range(len(masks)))                               # every gate exactly once
    assert [pos[wire[w]] for w in range(n)] == list(range(n))           # every wire's data is home
    return sum(1 for e in sched if e[0] == "block"), n_quads


def test_rotating_block_schedule_keeps_order_and_comes_home():
    assert _check_rotating_schedule(_brick_masks(20, 12), 20)[0] <= 24
    assert _check_rotating_schedule(_brick_masks(28, 100), 28)[0] <= 150     # 4150 gates (in place: 166 blocks)
    for seed in range(3):
        _check_rotating_schedule(_random_masks(19, 150, seed), 19)
        _check_rotating_schedule(_random_masks(26, 300, seed + 10), 26)


def _qft_program(n):
    masks, diag, swaps = [], [], []
    for top in range(n):
        masks.append(1 << top); diag.append(False); swaps.append(None)
        for w in range(top + 1, n):
            masks.append((1 << top) | (1 << w)); diag.append(True); swaps.append(None)
    for w in range(n // 2):
        masks.append((1 << w) | (1 << (n - 1 - w))); diag.append(False); swaps.append((w, n - 1 - w))
    return masks, diag, swaps


def test_qft_becomes_dense_blocks_and_quadratic_phase_passes():
    """Cooley-Tukey shape: a 6-qubit QFT block per window, the controlled phases between the windows as ONE
    quadratic-phase pass each, the final SWAPs as relabelling; the sort passes bring the wires home."""
    for n in (20, 31):
        masks, diag, swaps = _qft_program(n)
        blocks, quads = _check_rotating_schedule(masks, n, diag, swaps)
        # one block per 6-qubit window, one quadratic-phase pass per window boundary that carries >= QUAD_MIN phases;
        # the rest is layout: one parking pass at the start, <= 2 + ceil((n - 6) / 6) passes to bring the wires home
        assert quads >= (n - 8) // 6 and blocks + quads <= 2 * -(-n // 6) + 9, (n, blocks, quads)
    # random mixes of diagonal / non-diagonal gates and SWAPs keep their dependencies
    for seed in range(4):
        rng = np.random.default_rng(seed)
        n = 20
        masks, diag, swaps = [], [], []
        for _ in range(400):
            kind = rng.integers(4)
            qs = [int(q) for q in rng.choice(n, size=2, replace=False)]
            if kind == 0:
                masks.append(1 << qs[0]); diag.append(False); swaps.append(None)
            elif kind == 3 and rng.random() < 0.3:
                masks.append((1 << qs[0]) | (1 << qs[1])); diag.append(False); swaps.append((qs[0], qs[1]))
            elif kind >= 2:
                masks.append((1 << qs[0]) | (1 << qs[1])); diag.append(True); swaps.append(None)
            else:
                masks.append((1 << qs[0]) | (1 << qs[1])); diag.append(False); swaps.append(None)
        _check_rotating_schedule(masks, n, diag, swaps)


def test_block_stage_picks_the_cheaper_write_back():
    from torchquantum_b200 import blockplan as B

    deep, n_deep, _ = B.plan_blocks(_brick_masks(28, 100), 28)
    assert all(e[3] & C.KBF_ROTATE for e in deep)                        # deep circuits: rotating, fewer + cheaper passes
    assert B.schedule_rotating(_brick_masks(16, 8), 16) is None          # small states stay with the in-place schedule
    small, _, _ = B.plan_blocks(_brick_masks(16, 8), 16)
    assert not any(e[3] & C.KBF_ROTATE for e in small)


def test_rotating_schedule_can_leave_its_layout_to_the_caller():
    """final=[]: no passes that bring the qubits home; the logical -> physical map the stage ends with is reported (the
    sharded device composes it into its own layout) and equals the replay of the kernel's contract"""
    from torchquantum_b200 import blockplan as B

    n = 22
    masks = _brick_masks(n, 10)
    home = B.schedule_rotating(masks, n)
    final = []
    kept = B.schedule_rotating(masks, n, final=final)
    assert len(final) == 1 and sorted(final[0]) == list(range(n))
    assert len(kept) < len(home) and all(e[2] for e in kept)          # only blocks that carry gates
    pos = list(range(n))
    for _, kpos, gates, tr, where, perm in kept:
        pinv = {p: d for d, p in enumerate(pos)}
        arrive = list(kpos)
        for i, j in perm:
            arrive[i], arrive[j] = arrive[j], arrive[i]
        new = {}
        for i, p in enumerate(arrive):
            new[pinv[p]] = i if tr else 6 + i
        for b in range(6):
            new[pinv[b]] = 6 + b if tr else b
        r = 12
        for p in range(6, n):
            if p not in kpos:
                new[pinv[p]] = r
                r += 1
        pos = [new[d] for d in range(n)]
    assert pos == final[0]
