"""Shared test helpers: golden-fixture loading and tolerance definitions."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: 1e-5 relative on states / expectation values / gradients (complex64).
# "relative" = rel-L2 plus max-abs scaled by the largest reference magnitude.
RTOL = 1e-5
# gradients: the same bar (north_star); the reference's own tests use 1e-4 (test/static_mode_test.py)
GRAD_RTOL = 1e-5


def load(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    manifest = json.loads(str(z["manifest"]))
    return z, manifest


def assert_close(got, want, rtol=RTOL, what=""):
    got = torch.as_tensor(np.asarray(got) if not isinstance(got, torch.Tensor) else got.detach().cpu())
    want = torch.as_tensor(np.asarray(want) if not isinstance(want, torch.Tensor) else want.detach().cpu())
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    diff = (got - want).abs()
    scale = max(want.abs().max().item(), 1e-30)
    ref_norm = max(torch.linalg.vector_norm(want).item(), 1e-30)
    rel_l2 = torch.linalg.vector_norm(got - want).item() / ref_norm
    max_abs = diff.max().item() if diff.numel() else 0.0
    # absolute floor ONLY for small tensors of scalars that are analytically ~0 (cancelling expectation values /
    # gradients: rounding noise of O(eps) has no reference magnitude to be relative to); states never use it
    ok = (rel_l2 <= rtol and max_abs <= rtol * scale) or (got.numel() <= 64 and max_abs <= 2e-6)
    assert ok, f"{what}: rel_l2={rel_l2:.3e} max_abs={max_abs:.3e} scale={scale:.3e}"


def op_params(op, requires_grad=False):
    p = op.get("params")
    if p is None:
        return None
    t = torch.tensor(p, dtype=torch.float32)
    if requires_grad:
        t.requires_grad_(True)
    return t
