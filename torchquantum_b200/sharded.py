"""Sharded statevector device (north_star item 4; SURVEY §8 row (e), kernels K10 / K11).

The reference has no multi-GPU code at all (README.md:52 only claims "30+ qubits with multiple
GPUs"); this is the drop-in's extension of `tq.QuantumDevice` (device/devices.py:36-136) to states
that do not fit one GPU.  One process per GPU (`torch.distributed`, NCCL over NVLink 5):

* rank r of 2^g holds the amplitudes whose g top *physical* index bits equal r: a contiguous
  slab of 2^(n-g) amplitudes per batch element (TorchQuantum is big-endian, so with the initial
  layout these are wires 0..g-1);
* a table `pos[logical bit] -> physical bit` tracks where every qubit currently lives;
* gates whose targets are local run through the same fused gate-pass kernels as on one GPU;
  controls and parity bits of diagonal gates may sit in the rank index at no cost (the rank
  supplies the missing index bits to the kernel: K11, no traffic);
* a gate with a non-diagonal target on a global qubit triggers ONE exchange of all g global
  qubits with g local qubits that are not needed for the longest time.  Default back end ("p2p"): the
  slabs live in symmetric memory and ONE kernel per rank (`tqb_exchange_p2p`) stores every amplitude
  straight into the slab of the rank that owns it afterwards (NVLink peer stores; victims at any local
  positions, no pack pass).  Fallback ("collective", also what the gloo tests run): victims are moved to
  the top local bits by a fused local pass, then `all_to_all_single` of 2^g equal chunks swaps physical
  bits [nl-g, nl) <-> [nl, n).  Either way every rank sends (1 - 2^-g) of its shard;
* gates commute when they share no qubit, so before paying for an exchange the queue is scanned
  for everything that can still run in the current layout (same dependency-respecting greedy
  as the pass packer in plan.py);
* SWAP gates are free: they only relabel `pos`.

Autograd (north_star items 2 + 4 together): a flush issued with gradients enabled runs every local stage through the
same adjoint-differentiation nodes as on one GPU, and every exchange through `_ExchangeFn`, whose backward is the same
exchange applied to the incoming gradient (the swap of rank bits and local bits is an involution).  Expectation values
are summed over the ranks by an autograd-aware all-reduce.  Parameters shared by the ranks receive, on every rank, the
PARTIAL gradient of that rank's slab: `allreduce_gradients(params)` sums them (as a data-parallel trainer would).  In this
mode exchanges use the collective back end and each fused stage keeps its output slab for the backward.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _native, engine
from . import gates as G
from .device import GateQueue
from .engine import STATS
from .macro import C_DTYPE

__all__ = ["ShardedQuantumDevice"]


def _local_bits_needed(name, wires, params):
    """Wires (as given) that must be local for the gate to run without communication."""
    spec = G.SPECS[name]
    if spec.kind in (G.DIAG, G.NOP):
        return ()
    if spec.kind == G.RESET:
        raise NotImplementedError("reset is not supported on a sharded device")
    if spec.kind == G.DENSE:
        return tuple(wires)
    n_ctrl = spec.n_ctrl if spec.n_ctrl >= 0 else len(wires) - 1
    return tuple(wires[n_ctrl:])


class _ExchangeFn(torch.autograd.Function):
    """Swap physical bits [nl-g, nl) <-> [nl, n) with all_to_all_single; the backward is the same swap of the gradient."""

    @staticmethod
    def forward(ctx, flat, dev):
        ctx.dev = dev
        return dev._collective_swap(flat)

    @staticmethod
    def backward(ctx, g):
        return ctx.dev._collective_swap(g.contiguous()), None


class _AllReduceSumFn(torch.autograd.Function):
    """sum over the ranks; d(sum)/d(local) = 1, so the gradient passes through unchanged"""

    @staticmethod
    def forward(ctx, t, dev):
        return dev._all_reduce(t.clone())

    @staticmethod
    def backward(ctx, g):
        return g, None


class ShardedQuantumDevice:
    """`QuantumDevice` whose statevector is partitioned over the ranks of a process group."""

    is_sharded = True

    def __init__(self, n_wires, device_name="default", bsz=1, device="cuda", record_op=False, process_group=None,
                 exchange="auto"):
        """exchange: "p2p" = pack + transfer in one kernel that stores straight into the peers' slabs over NVLink
        (symmetric memory; CUDA only), "collective" = local permute pass + torch.distributed all_to_all_single (NCCL on
        GPUs, gloo in the CPU tests), "auto" = p2p when the symmetric-memory rendezvous succeeds."""
        if not dist.is_initialized():
            raise RuntimeError("ShardedQuantumDevice needs torch.distributed to be initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.g = int(np.log2(self.world))
        if 1 << self.g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n_wires, self.device_name, self.bsz = n_wires, device_name, bsz
        self.nl = n_wires - self.g
        if self.nl < 1 or (self.g and self.nl < max(2 * self.g - 1, self.g + 2)):
            # an exchange parks g local qubits in the rank index while a 2-qubit gate keeps 2 on chip, and
            # canonicalize() parks g more beside the g - 1 it is sorting
            raise ValueError(f"too few qubits ({n_wires}) for {self.world} ranks: need n - g >= max(2g - 1, g + 2)")
        self.device = torch.device(device)
        self.record_op = record_op
        self._op_history, self._history_pending = [], []
        self._queue = GateQueue(self)
        self.stats = {"exchanges": 0, "exchange_bytes_sent": 0, "exchange_ms": [], "local_permute_passes": 0}
        self.time_exchanges = False
        if exchange not in ("auto", "p2p", "collective"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'collective'")
        self._exchange_pref = exchange
        self._symm = None   # (two symmetric slabs, their rendezvous handles, index of the slab holding the state)
        self.reset_states(bsz)

    # --- state container ------------------------------------------------------------------
    def _new_slab(self, bsz):
        """The slab this rank's state lives in.  With the p2p exchange it is one of two symmetric-memory buffers (every
        rank can store into every other rank's copy); otherwise a plain device tensor."""
        want = self._exchange_pref != "collective" and self.device.type == "cuda" and self.g > 0
        if want and (self._symm is None or self._symm[0][0].numel() != 2 * (bsz << self.nl)):
            try:
                import torch.distributed._symmetric_memory as symm_mem

                group = self.group if self.group is not None else dist.group.WORLD
                bufs = [symm_mem.empty(2 * (bsz << self.nl), dtype=torch.float32, device=self.device) for _ in range(2)]
                handles = [symm_mem.rendezvous(b, group) for b in bufs]
                self._symm = [bufs, handles, 0]
            except Exception as e:  # noqa: BLE001  (no peer access, old driver, unsupported backend ...)
                if self._exchange_pref == "p2p":
                    raise RuntimeError(f"ShardedQuantumDevice(exchange='p2p'): symmetric memory is not available: {e}") from e
                self._symm = None
                self._exchange_pref = "collective"
        if not want or self._symm is None:
            self._symm = None
            return torch.empty(bsz, 1 << self.nl, dtype=C_DTYPE, device=self.device)
        self._symm[2] = 0
        return self._symm_view(0, bsz)

    def _symm_view(self, which, bsz):
        return torch.view_as_complex(self._symm[0][which].view(bsz, 1 << self.nl, 2))

    @property
    def exchange_backend(self):
        return "p2p" if self._symm is not None else "collective"

    def reset_states(self, bsz):
        self.bsz = bsz
        flat = self._new_slab(bsz)
        flat.zero_()
        if self.rank == 0:
            flat[:, 0] = 1
        self._flat = flat
        self._queue = GateQueue(self)
        self.pos = list(range(self.n_wires))  # logical bit -> physical bit

    def set_local_states(self, flat):
        """Install this rank's slab (canonical layout: global index = rank * 2^nl + local index)."""
        flat = flat.reshape(flat.shape[0], 1 << self.nl)
        self.bsz = flat.shape[0]
        slab = self._new_slab(self.bsz)
        slab.copy_(flat)
        self._flat = slab
        self._queue = GateQueue(self)
        self.pos = list(range(self.n_wires))

    @property
    def compute_device(self):
        return self._flat.device

    def enqueue(self, name, wires, params, inverse):
        self._queue.append((name, wires, params, inverse))

    def reset_op_history(self):
        self._op_history, self._history_pending = [], []

    @property
    def op_history(self):
        for entry in self._history_pending:
            p = entry["params"]
            if p is not None:
                entry["params"] = p.squeeze().detach().cpu().numpy().tolist()
        self._history_pending = []
        return self._op_history

    def _record(self, entry):
        self._op_history.append(entry)
        self._history_pending.append(entry)

    # --- execution ------------------------------------------------------------------------
    def _pwire(self, logical_bit, pos=None):
        return self.n_wires - 1 - (self.pos if pos is None else pos)[logical_bit]

    def _run_local(self, records, leave=False):
        """leave: the gates of the user's circuit (flush): rotating block stages may leave the local qubits in another
        layout, which goes into `pos`; the device's own permutation passes need the layout they ask for."""
        if records:
            timed = self.time_exchanges and self._flat.is_cuda
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = engine.STATS["launches"]
                e0.record()
            grad = getattr(self, "_grad_flush", False)
            spare = self._symm_view(1 - self._symm[2], self.bsz) if (self._symm is not None and not grad) else None
            out = engine.run_circuit(self._flat, self.nl, records, n_total=self.n_wires, rank=self.rank,
                                     inplace=not grad, scratch=spare, leave_layout=leave and not grad)
            if leave and not grad:
                # rotating block stages do not sort the local qubits home: their final layout goes into `pos`
                out, moved = out
                if moved is not None:
                    self.pos = [moved[p] if p < self.nl else p for p in self.pos]
                    self.stats["layouts_kept"] = self.stats.get("layouts_kept", 0) + 1
            if grad:
                self._flat = out          # autograd: every stage is out of place and keeps its output for the backward
            elif spare is not None and out.data_ptr() == spare.data_ptr():
                self._symm[2] = 1 - self._symm[2]   # rotating block passes left the state in the other symmetric slab
                self._flat = spare
            elif self._symm is not None and out.data_ptr() != self._flat.data_ptr():
                self._flat.copy_(out)   # an out-of-place stage (dense block, reset): back into the symmetric slab
            else:
                self._flat = out
            if timed:
                e1.record()
                self.stats.setdefault("local_runs", []).append((len(records), engine.STATS["launches"] - l0, e0, e1))

    def flush(self):
        if not self._queue:
            return
        remaining, self._queue = self._queue, GateQueue(self)
        # same call-time semantics as QuantumDevice (device.GateQueue): parameters must not have been edited in place
        # since the gate call, and the gates run under the autograd mode they were ISSUED under
        remaining.check()
        n, nl = self.n_wires, self.nl
        grad = remaining.grad and (self._flat.requires_grad or any(
            r[2] is not None and r[2].requires_grad for r in remaining))
        self._grad_flush = grad
        if not grad and self._flat.requires_grad:
            # gates issued under no_grad on a state that carries a graph: the reference replaces the state tensor and
            # leaves the old one to its graph; the in-place passes here must not write into a slab a backward still needs
            self._flat = self._own_copy(self._flat)
        with torch.set_grad_enabled(grad):
            while remaining:
                runnable, rest, blocked = [], [], 0
                pos = list(self.pos)
                for rec in remaining:
                    name, wires, params, inverse = rec
                    lbits = [n - 1 - w for w in wires]
                    qm = 0
                    for b in lbits:
                        qm |= 1 << b
                    need = _local_bits_needed(name, wires, params)
                    if name == "swap" and not (qm & blocked):
                        pos[lbits[0]], pos[lbits[1]] = pos[lbits[1]], pos[lbits[0]]  # relabel: no data moves
                        continue
                    if (qm & blocked) or any(pos[n - 1 - w] >= nl for w in need):
                        blocked |= qm
                        rest.append(rec)
                        continue
                    runnable.append((name, tuple(n - 1 - pos[b] for b in lbits), params, inverse))
                self.pos = pos
                self._run_local(runnable, leave=True)
                if rest:
                    self._exchange(rest)
                remaining = rest

    def _own_copy(self, flat):
        """a graph-free copy of the slab that in-place passes may overwrite (in the symmetric slab when there is one)"""
        if self._symm is None:
            return flat.detach().clone()
        slab = self._symm_view(self._symm[2], self.bsz)
        if slab.data_ptr() != flat.data_ptr():
            slab.copy_(flat.detach())
        return slab

    def _next_need(self, rest):
        """logical bit -> index of the first record of `rest` that needs it local (inf if none)."""
        nxt = {}
        for i, (name, wires, params, inverse) in enumerate(rest):
            for w in _local_bits_needed(name, wires, params):
                nxt.setdefault(self.n_wires - 1 - w, i)
        return nxt

    def _exchange(self, rest):
        """Bring every global qubit on chip: pick the g local qubits needed last as victims, move them
        to the top local bits (fused local pass), all-to-all."""
        n, nl, g = self.n_wires, self.nl, self.g
        nxt = self._next_need(rest)
        name0, wires0, params0, _ = rest[0]
        need0 = {n - 1 - w for w in _local_bits_needed(name0, wires0, params0)}
        if len(need0) > nl - g:
            raise NotImplementedError(
                f"{name0} on wires {list(wires0)} needs {len(need0)} qubits on one GPU at once, but with {self.world} "
                f"ranks only {nl - g} of the {nl} local qubits can stay while the {g} rank qubits are exchanged: "
                "decompose the gate (e.g. tq.QFT / QFTLayer instead of the dense tqf.qft) or use fewer ranks")
        inv = {p: l for l, p in enumerate(self.pos)}
        local_logical = [inv[p] for p in range(nl) if inv[p] not in need0]
        if self._symm is not None:
            # a victim at physical bit p leaves runs of 2^p amplitudes contiguous on both sides of the transfer: below bit 6
            # the peer stores degrade to 8-256 byte pieces (30 ms instead of 14 ms per exchange at n = 32 on 2 GPUs), so
            # low positions are only taken when nothing else is left
            high = [l for l in local_logical if self.pos[l] >= 6]
            if len(high) >= g:
                local_logical = high
        order = sorted(local_logical, key=lambda l: (-nxt.get(l, 1 << 60), -self.pos[l]))
        victims = order[:g]
        if self._symm is not None and not getattr(self, "_grad_flush", False):
            self._exchange_p2p(sorted(self.pos[v] for v in victims))   # no pack pass: any local positions
        else:
            self._move_to_top(victims)
            self._all_to_all()
        if any(self.pos[l] >= nl for l in need0):   # progress guarantee of flush()
            raise AssertionError(f"exchange left {name0} on wires {list(wires0)} blocked (layout {self.pos})")

    def _move_to_top(self, victims):
        """Local bit permutation so that `victims` (logical bits) occupy physical bits [nl-g, nl)."""
        n, nl, g = self.n_wires, self.nl, self.g
        top = list(range(nl - g, nl))
        inv = {p: l for l, p in enumerate(self.pos)}
        vict = set(victims)
        free_top = [p for p in top if inv[p] not in vict]
        movers = [v for v in victims if self.pos[v] < nl - g]
        recs = []
        for v, p in zip(movers, free_top):
            a, b = self.pos[v], p                      # exchange the contents of physical bits a and b
            wa, wb = n - 1 - a, n - 1 - b
            recs += [("cnot", (wa, wb), None, False), ("cnot", (wb, wa), None, False), ("cnot", (wa, wb), None, False)]
            other = inv[p]
            self.pos[v], self.pos[other] = b, a
            inv[a], inv[b] = other, v
        if recs:
            self.stats["local_permute_passes"] += 1
            self._run_local(recs)

    def _exchange_p2p(self, vpos):
        """Swap rank-index bit k <-> local physical bit vpos[k] (ascending) with ONE kernel per rank that stores every
        amplitude straight into the slab of the rank that owns it afterwards (NVLink peer stores into symmetric memory),
        then a device-side barrier over the symmetric-memory signal pads.  K10, fused pack + exchange."""
        nl, g = self.nl, self.g
        bufs, handles, cur = self._symm
        dst = 1 - cur
        timed = self.time_exchanges
        # the destination slab is every rank's spare one, which rotating block passes use as their second buffer: nobody
        # may store into a peer's spare slab before that peer's local passes are done
        handles[dst].barrier(channel=1)
        if timed:   # from "every rank is ready" to "every rank's data has landed"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _native.call("tqb_exchange_p2p", self._flat, bufs[dst], [int(p) for p in handles[dst].buffer_ptrs], nl, self.bsz,
                     [int(p) for p in vpos], self.rank)
        STATS["launches"] += 1
        handles[dst].barrier(channel=0)   # every rank's stores have landed before anybody reads its new slab
        if timed:
            e1.record()
            self.stats["exchange_ms"].append((e0, e1))
        self.stats["exchanges"] += 1
        self.stats["exchange_bytes_sent"] += (self.world - 1) * (1 << (nl - g)) * self.bsz * 8
        self.stats.setdefault("exchange_victim_bits", []).append([int(p) for p in vpos])
        self._symm[2] = dst
        self._flat = self._symm_view(dst, self.bsz)
        where = {p: k for k, p in enumerate(vpos)}
        for l, p in enumerate(self.pos):
            if p in where:
                self.pos[l] = nl + where[p]
            elif p >= nl:
                self.pos[l] = vpos[p - nl]

    def _all_to_all(self):
        """Swap physical bits [nl-g, nl) <-> [nl, n): chunk j of the slab goes to rank j and the chunk
        received from rank i lands at chunk position i (K10)."""
        nl, g, W = self.nl, self.g, self.world
        grad = getattr(self, "_grad_flush", False) and torch.is_grad_enabled()
        if self._symm is not None and not grad:
            return self._exchange_p2p(list(range(nl - g, nl)))
        flat = self._flat
        bsz = flat.shape[0]
        chunk = 1 << (nl - g)
        timed = self.time_exchanges and flat.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._flat = _ExchangeFn.apply(flat, self) if grad else self._collective_swap(flat)
        if timed:
            e1.record()
            self.stats["exchange_ms"].append((e0, e1))
        self.stats["exchanges"] += 1
        self.stats["exchange_bytes_sent"] += (W - 1) * chunk * bsz * 8
        for l, p in enumerate(self.pos):
            if nl - g <= p < nl:
                self.pos[l] = p + g
            elif p >= nl:
                self.pos[l] = p - g

    def _collective_swap(self, flat):
        """the data movement of _all_to_all on an arbitrary [bsz, 2^nl] tensor (state or gradient)"""
        nl, g, W = self.nl, self.g, self.world
        flat = flat.contiguous()
        bsz = flat.shape[0]
        chunk = 1 << (nl - g)
        if bsz == 1:
            send = flat.view(W, chunk)
        else:
            send = flat.view(bsz, W, chunk).transpose(0, 1).contiguous()
        recv = torch.empty_like(send)
        if flat.is_cuda and dist.get_backend(self.group) == "gloo":
            # a process group that cannot move device memory (several ranks sharing one GPU in the test-suite): the
            # chunks go through pinned host buffers; the kernels on either side are the same
            h_send = torch.view_as_real(send).cpu()
            h_recv = torch.empty_like(h_send)
            dist.all_to_all_single(h_recv, h_send, group=self.group)
            torch.view_as_real(recv).copy_(h_recv)
        else:
            dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=self.group)
        if bsz == 1:
            return recv.view(1, -1)
        return recv.view(W, bsz, chunk).transpose(0, 1).contiguous().view(bsz, -1)

    def local_run_times(self):
        """[(gates, kernel launches, ms)] of every local fused run since the stats were reset (time_exchanges)."""
        return [(n, l, round(a.elapsed_time(b), 3)) for n, l, a, b in self.stats.get("local_runs", [])]

    def exchange_times_ms(self):
        out = [a.elapsed_time(b) for a, b in self.stats["exchange_ms"]]
        return out

    # --- layout ---------------------------------------------------------------------------
    def canonicalize(self):
        """Restore the identity layout (logical bit l at physical bit l)."""
        self.flush()
        n, nl, g = self.n_wires, self.nl, self.g
        # a state that carries a graph keeps it through the layout change, whatever mode the READ happens under (a
        # `no_grad` peek must neither cut the graph nor permute a slab the backward still needs in place)
        self._grad_flush = self._flat.requires_grad
        with torch.set_grad_enabled(self._grad_flush):
            target_global = list(range(nl, n))
            if any(self.pos[l] != l for l in target_global):
                if any(self.pos[l] >= nl for l in target_global):
                    # some rank qubits are in the rank index at the wrong place (or stay there): first
                    # bring ALL of them on chip, parking g qubits that belong on chip in the rank index
                    inv = {p: l for l, p in enumerate(self.pos)}
                    park = [inv[p] for p in range(nl) if inv[p] < nl][:g]
                    self._move_to_top(park)
                    self._all_to_all()
                self._move_to_top_ordered(target_global)
                self._all_to_all()
            self._permute_local_to_identity()
        assert self.pos == list(range(n))

    def _move_to_top_ordered(self, logical):
        """logical[i] -> physical bit nl-g+i (all of them currently local)."""
        n, nl, g = self.n_wires, self.nl, self.g
        self._apply_local_permutation({l: nl - g + i for i, l in enumerate(logical)})

    def _permute_local_to_identity(self):
        self._apply_local_permutation({l: l for l in range(self.nl)})

    def _apply_local_permutation(self, want):
        """want: logical bit -> physical local bit it must end at (other local qubits fill what is left)."""
        n, nl = self.n_wires, self.nl
        recs = []
        inv = {p: l for l, p in enumerate(self.pos)}
        for l, dst in want.items():
            a = self.pos[l]
            assert a < nl and dst < nl
            if a == dst:
                continue
            wa, wb = n - 1 - a, n - 1 - dst
            recs += [("cnot", (wa, wb), None, False), ("cnot", (wb, wa), None, False), ("cnot", (wa, wb), None, False)]
            other = inv[dst]
            self.pos[l], self.pos[other] = dst, a
            inv[a], inv[dst] = other, l
        if recs:
            self.stats["local_permute_passes"] += 1
            self._run_local(recs)

    # --- reading the state ----------------------------------------------------------------
    def get_states_1d(self):
        """This rank's slab [bsz, 2^(n-g)] in the canonical layout (amplitude rank * 2^nl + k)."""
        self.canonicalize()
        return self._flat

    def gather_states_1d(self):
        """Full [bsz, 2^n] state on every rank (tests / small n only)."""
        local = self.get_states_1d().contiguous()
        if local.is_cuda and dist.get_backend(self.group) == "gloo":
            h = torch.view_as_real(local).cpu()
            hp = [torch.empty_like(h) for _ in range(self.world)]
            dist.all_gather(hp, h, group=self.group)
            return torch.cat([torch.view_as_complex(x) for x in hp], dim=1).to(local.device)
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(local), group=self.group)
        return torch.cat(parts, dim=1)

    def read_amplitudes(self, indices):
        """Amplitudes at the given canonical flat indices (python ints) in whatever layout the state is
        in right now: the owner rank looks each one up, one all-reduce shares them -> [bsz, len]."""
        self.flush()
        n, nl = self.n_wires, self.nl
        out = torch.zeros(self.bsz, len(indices), 2, dtype=torch.float32, device=self._flat.device)
        for i, k in enumerate(indices):
            phys = 0
            for l in range(n):
                if (k >> l) & 1:
                    phys |= 1 << self.pos[l]
            if phys >> nl == self.rank:
                out[:, i] = torch.view_as_real(self._flat[:, phys & ((1 << nl) - 1)])
        return torch.view_as_complex(self._all_reduce(out))

    @property
    def states(self):
        raise RuntimeError("a sharded device exposes get_states_1d() (local slab) / gather_states_1d()")

    # --- measurement (dispatched from torchquantum_b200.measurement) -----------------------
    def _all_reduce(self, t):
        if t.is_cuda and dist.get_backend(self.group) == "gloo":
            h = t.cpu()
            dist.all_reduce(h, group=self.group)
            t.copy_(h)
            return t
        dist.all_reduce(t, group=self.group)
        return t

    def expval_pauli(self, observables):
        """[bsz, n_terms] <psi|P_t|psi>: X / Y factors on rank-index qubits are rotated to Z first
        (the rotation is undone lazily by queued inverse gates), Z factors there only flip the sign of
        this rank's partial sum; one all-reduce of bsz * n_terms floats."""
        from . import functional as tqf
        from .measurement import pauli_masks

        n, nl = self.n_wires, self.nl
        self.flush()
        out_cols = []
        # group terms that need no rotation
        direct, rotated = [], []
        for t, o in enumerate(observables):
            assert len(o) == n
            needs = any(ch in "XY" and self.pos[n - 1 - w] >= nl for w, ch in enumerate(o.upper()))
            (rotated if needs else direct).append(t)
        result = torch.zeros(self.bsz, len(observables), dtype=torch.float32, device=self._flat.device)
        if direct:
            result[:, direct] = self._expval_local([observables[t] for t in direct])
        for t in rotated:
            o = observables[t].upper()
            undo = []
            zs = []
            for w, ch in enumerate(o):
                if ch == "X":
                    tqf.hadamard(self, wires=w)
                    undo.append(("hadamard", w))
                elif ch == "Y":
                    tqf.sdg(self, wires=w)
                    tqf.hadamard(self, wires=w)
                    undo += [("s", w), ("hadamard", w)]
                zs.append("I" if ch == "I" else "Z")
            self.flush()
            result[:, t] = self._expval_local(["".join(zs)])[:, 0]
            for name, w in reversed(undo):
                getattr(tqf, name)(self, wires=w)
        if torch.is_grad_enabled() and result.requires_grad:
            return _AllReduceSumFn.apply(result, self)
        return self._all_reduce(result)

    def allreduce_gradients(self, params):
        """Sum the `.grad` of parameters shared by the ranks (each rank's backward yields the partial gradient of its
        slab), in place: call after `loss.backward()`, as a data-parallel trainer would."""
        grads = [p.grad for p in params if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._all_reduce(flat)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def _expval_local(self, observables):
        from .measurement import pauli_masks

        n, nl = self.n_wires, self.nl
        xs, zs, nys, signs = [], [], [], []
        for o in observables:
            x = z = ny = 0
            sign = 1.0
            for w, ch in enumerate(o.upper()):
                p = self.pos[n - 1 - w]
                if ch == "I":
                    continue
                if p >= nl:
                    assert ch == "Z"
                    if (self.rank >> (p - nl)) & 1:
                        sign = -sign
                    continue
                if ch in "XY":
                    x |= 1 << p
                if ch in "YZ":
                    z |= 1 << p
                if ch == "Y":
                    ny += 1
            xs.append(x); zs.append(z); nys.append(ny); signs.append(sign)
        dev = self._flat.device
        xm = torch.tensor(xs, dtype=torch.int64, device=dev)
        zm = torch.tensor(zs, dtype=torch.int64, device=dev)
        ny = torch.tensor(nys, dtype=torch.int32, device=dev)
        state = self._flat
        _native.check_device(state)
        if torch.is_grad_enabled() and state.requires_grad:
            from .measurement import _ExpvalPauliFn

            out = _ExpvalPauliFn.apply(state, nl, xm, zm, ny, tuple(observables), None, None)
        else:
            out = torch.zeros(self.bsz, len(observables), dtype=torch.float32, device=dev)
            _native.call("tqb_expval_pauli", state, nl, self.bsz, xm, zm, ny, len(observables), out)
            STATS["launches"] += 1
        return out * torch.tensor(signs, dtype=torch.float32, device=dev)

    def expval_z_all(self):
        """[bsz, n+1]: column l = <Z> of LOGICAL bit l, column n = squared norm (all-reduced)."""
        n, nl = self.n_wires, self.nl
        self.flush()
        state = self._flat
        if torch.is_grad_enabled() and state.requires_grad:
            raise NotImplementedError("per-wire expval / MeasureAll on a sharded device is not differentiable: "
                                      "use expval_joint_analytical (e.g. 'IIZI') for a loss, or read under torch.no_grad()")
        dev = state.device
        loc = torch.zeros(self.bsz, nl + 1, dtype=torch.float32, device=dev)
        _native.call("tqb_expval_z", state, nl, self.bsz, loc)
        STATS["launches"] += 1
        out = torch.empty(self.bsz, n + 1, dtype=torch.float32, device=dev)
        norm = loc[:, nl]
        out[:, n] = norm
        for l in range(n):
            p = self.pos[l]
            if p < nl:
                out[:, l] = loc[:, p]
            else:
                out[:, l] = -norm if (self.rank >> (p - nl)) & 1 else norm
        return self._all_reduce(out)

    @property
    def name(self):
        return self.__class__.__name__

    def __repr__(self):
        return (f" class: {self.name} \n number of qubits: {self.n_wires} ({self.nl} local + {self.g} in the rank index)\n"
                f" batch size: {self.bsz} \n rank {self.rank} of {self.world} \n layout: {self.pos}")


def _bind_functionals():
    from .functional import func_name_dict, func_name_dict_collect

    for name, func in func_name_dict.items():
        setattr(ShardedQuantumDevice, name, func)
    for name, func in func_name_dict_collect.items():
        setattr(ShardedQuantumDevice, name, func)


_bind_functionals()
