"""Differential checks against the LIVE reference (only where /root/reference is mounted: the build container, not the
GPU box): the same calls on mit-han-lab/torchquantum (through oracle/refshim.py) and on this package (host emulation of
the kernel source) must give the same values, and where the reference raises, this package must not return silently
different results.  The scripts are tools/diff_api.py (device / functional / measurement / module API incl. error
behaviour), tools/diff_ops.py (every class of `op_name_dict`) and tools/diff_layers.py (layer templates with their
constructor options, parameter initialisation included)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/torchquantum"), reason="reference tree not mounted")


def _run(tool):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    problems = [l for l in lines if "VALUE-DIFF" in l or "ours-err" in l]
    # "[DIFF] name: ref=err ... ours=ok": the reference's own code is broken there (e.g. SingleExcitation has no
    # build_params at this commit); the other direction — the reference works and this package raises — is a failure
    problems += [l for l in lines if l.startswith("[DIFF]") and "ours=ok" not in l]
    assert not problems, "\n".join(problems)
    return lines


def test_api_calls_match_the_live_reference():
    lines = _run("diff_api.py")
    assert sum(l.startswith("[ok]") for l in lines) >= 35


def test_every_operator_class_matches_the_live_reference():
    lines = _run("diff_ops.py")
    assert lines[-1].startswith("ok 72 bad 3 "), lines[-1]     # 3 = classes that only work here (reference bugs), listed as [DIFF] ... ours=ok


def test_layer_templates_match_the_live_reference():
    lines = _run("diff_layers.py")
    assert lines[-1].strip() == "ok 13 bad 0", lines[-1]
