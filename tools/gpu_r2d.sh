#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe.json > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"; grep -v "^\[" gpurun_out/umma_probe.log | tail -12; grep -c "ERROR\|differs" gpurun_out/umma_probe.log; sort -t' ' -k8 -g gpurun_out/umma_probe.log | grep max_rel | tail -2
for d in 2 4 8 12; do
  echo "== TQB200_BLOCK_DEBUG=$d"
  TQB200_BLOCK_DEBUG=$d timeout 120 python tools/umma_probe.py --n 0 --time-n 28 --out gpurun_out/umma_dbg_$d.json 2>&1 | grep -v "^norm"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
