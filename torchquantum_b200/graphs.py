"""CUDA-graph capture of a whole training step.

A step of a small-qubit batched workload (config 1/2) is a few dozen short kernels; the Python that
queues 600 gates, walks autograd and accumulates 420 parameter gradients costs more than the kernels
run.  `GraphedStep` records one forward + backward of `loss_fn` into a CUDA graph (the plan, the gate
programs and the Pauli masks are already resident after the warm-up steps) and replays it per step:
inputs are copied into the captured buffers, the loss and the parameters' `.grad` tensors are the
captured outputs (and `static_inputs[i].grad` for inputs that require grad).  Gate parameters are read from the live `nn.Parameter` storage at replay time, so an
optimizer stepping them in place between replays works as usual.

The circuit structure, batch size and parameter set must not change between replays (the same
restriction as `torch.cuda.make_graphed_callables`)."""
import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, loss_fn, example_inputs, params, warmup=3):
        self.params = [p for p in params if p.requires_grad]
        self.static_inputs = [x.detach().clone().requires_grad_(x.requires_grad) for x in example_inputs]
        dev = self.static_inputs[0].device
        assert dev.type == "cuda", "GraphedStep needs CUDA tensors"
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self._drop_stale_graphs()
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._zero()
                loss_fn(*self.static_inputs).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._zero()
        self._drop_stale_graphs()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn(*self.static_inputs)
            self.loss.backward()
            # all parameter gradients as ONE tensor, produced by the replay itself: data-parallel training all-reduces
            # this (one collective, no per-parameter Python) and `allreduce_gradients` writes the sums back
            pg = [p.grad for p in self.params if p.grad is not None]
            self.flat_grad = torch.cat([g.reshape(-1) for g in pg]) if pg else None
        self._leaves = self.params + [x for x in self.static_inputs if x.requires_grad]
        self._grads = [p.grad for p in self._leaves]   # the captured gradient buffers
        self._scatter = None
        if self.flat_grad is not None:
            self._scatter = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._scatter):
                torch._foreach_copy_([g.view(-1) for g in pg], list(self.flat_grad.split([g.numel() for g in pg])))

    @staticmethod
    def _drop_stale_graphs():
        """The read-out cache (measurement._READOUT_CURRENT) holds the columns of the last state that was read, and through
        their grad_fn the autograd graph of that step with the parameters' AccumulateGrad nodes.  A node that survives
        is reused by the next step on the stream it was created on, which breaks the capture (the warm-up and the
        capture run on their own streams): drop it so that every phase builds fresh nodes."""
        from . import measurement

        measurement._READOUT_CURRENT.clear()

    def _zero(self):
        for p in self.params + self.static_inputs:
            p.grad = None

    def allreduce_gradients(self, group=None, average=False):
        """Sum (or average) the parameter gradients of the last replay over the ranks of `group`: ONE all-reduce of the
        flat gradient the replay produced, then a captured scatter back into every parameter's `.grad`."""
        import torch.distributed as dist

        if self.flat_grad is None:
            return
        dist.all_reduce(self.flat_grad, group=group)
        if average:
            self.flat_grad /= dist.get_world_size(group)
        self._scatter.replay()

    def __call__(self, *inputs):
        """Copy `inputs` (device or pinned host tensors) into the captured buffers, replay the step and
        return the captured loss tensor; gradients are in the parameters' `.grad`."""
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"GraphedStep was captured with {len(self.static_inputs)} inputs, got {len(inputs)}")
        with torch.no_grad():
            for s, x in zip(self.static_inputs, inputs):
                if x.shape != s.shape:   # (copy_ would broadcast a smaller batch silently)
                    raise ValueError(f"GraphedStep was captured for an input of shape {tuple(s.shape)}, got "
                                     f"{tuple(x.shape)}: the batch size and circuit structure are fixed at capture")
                if x is not s:
                    s.copy_(x, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self._leaves, self._grads):   # survive `zero_grad(set_to_none=True)` between steps
            if p.grad is not g:
                p.grad = g
        return self.loss
