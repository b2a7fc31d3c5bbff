#!/bin/bash
# rotating block write-back: numerics + timing of the kernel, planned-circuit timings for both disciplines, depth parity
mkdir -p gpurun_out
timeout 600 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe_rot.json > gpurun_out/umma_probe_rot.log 2>&1
echo "probe rc=$?"; grep -c "rel_l2" gpurun_out/umma_probe_rot.log; grep -i "error\|differs" gpurun_out/umma_probe_rot.log | head; grep "rotate\|unordered" gpurun_out/umma_probe_rot.log | awk '{print $NF}' | sort -g | tail -2
grep "block\|rot_\|three" gpurun_out/umma_probe_rot.log | tail -16
for m in inplace rotate; do
  TQB200_BLOCK_WRITEBACK=$m timeout 300 python tools/block_times.py --n 28 --layers 16 > gpurun_out/block_times_$m.txt 2>&1
  echo "$m: $(tail -1 gpurun_out/block_times_$m.txt)"
done
TQB200_BLOCK_WRITEBACK=rotate timeout 900 python -m pytest tests -m gpu -x -q -k "depth100 or cfg4_full" 2>&1 | tail -3
for m in inplace rotate auto; do
  TQB200_BLOCK_WRITEBACK=$m timeout 600 python tools/configs.py --cfg 4 --out gpurun_out/cfg4_$m.json 2>&1 | tail -2
done
