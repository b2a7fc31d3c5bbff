#!/bin/bash
# raw-hi pipeline vs split-then-multiply: numerics, burst timing, sustained timing (depth-100 plan), parity
mkdir -p gpurun_out
timeout 600 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe_rawhi.json > gpurun_out/umma_probe_rawhi.log 2>&1
echo "probe rc=$?"; grep -i "error\|differs" gpurun_out/umma_probe_rawhi.log | head -5; grep "rel_l2" gpurun_out/umma_probe_rawhi.log | awk '{print $NF}' | sort -g | tail -2; grep "norm drift\|_block\|rot_" gpurun_out/umma_probe_rawhi.log | tail -18
TQB200_BLOCK_WRITEBACK=rotate timeout 300 python tools/block_times.py --n 28 --layers 100 > gpurun_out/block_times_rawhi_d100.txt 2>&1; echo "rawhi d100: $(tail -1 gpurun_out/block_times_rawhi_d100.txt)"
cp torchquantum_b200/lib/libtqb200.so /tmp/keep.so; cp tools/_ab/libtqb200_presplit.so torchquantum_b200/lib/libtqb200.so
TQB200_BLOCK_WRITEBACK=rotate timeout 300 python tools/block_times.py --n 28 --layers 100 > gpurun_out/block_times_presplit_d100.txt 2>&1; echo "presplit d100: $(tail -1 gpurun_out/block_times_presplit_d100.txt)"
cp /tmp/keep.so torchquantum_b200/lib/libtqb200.so
TQB200_BLOCK_WRITEBACK=rotate timeout 300 python tools/block_times.py --n 28 --layers 100 > gpurun_out/block_times_rawhi_d100b.txt 2>&1; echo "rawhi d100 again: $(tail -1 gpurun_out/block_times_rawhi_d100b.txt)"
timeout 900 python -m pytest tests -m gpu -x -q -k "depth100 or cfg4_full or write_back" 2>&1 | tail -3
timeout 600 python tools/configs.py --cfg 4 --out gpurun_out/cfg4_rawhi.json 2>&1 | tail -1
