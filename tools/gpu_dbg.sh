#!/bin/bash
# block-kernel stage ablation (timing only; results are garbage for debug != 0)
mkdir -p gpurun_out
for d in 0 1 2 4 8 3 5 6 7 12 15; do
  echo "== TQB200_BLOCK_DEBUG=$d"
  TQB200_BLOCK_DEBUG=$d timeout 120 python tools/umma_probe.py --n 0 --time-n 28 --out gpurun_out/umma_dbg_$d.json 2>&1 | grep -v "^norm"
done
