#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "sharded" 2>&1 | tail -4
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python - > gpurun_out/sanitizer_gemm.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from torchquantum_b200 import engine
torch.manual_seed(0)
for M, K, N in ((300, 128, 128), (128, 256, 512), (256, 1024, 128)):
    a = torch.randn(M, K, dtype=torch.complex64, device="cuda") / K ** 0.5
    b = torch.randn(K, N, dtype=torch.complex64, device="cuda")
    got = engine.complex_gemm(a, b)
    want = a.to(torch.complex128) @ b.to(torch.complex128)
    print(M, K, N, ((got - want).norm() / want.norm()).item())
PY
echo "sanitizer gemm rc=$?"; tail -6 gpurun_out/sanitizer_gemm.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/umma_probe.py --n 14 --time-n 14 --out gpurun_out/x.json > gpurun_out/sanitizer_block.log 2>&1; echo "sanitizer block rc=$?"; tail -4 gpurun_out/sanitizer_block.log
