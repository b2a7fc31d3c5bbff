#!/bin/bash
mkdir -p gpurun_out
python tools/cfg2_live_kernels.py > gpurun_out/cfg2_live_kernels.txt 2>&1; tail -30 gpurun_out/cfg2_live_kernels.txt
python tools/block_trace.py --n 28 --kpos 20 21 22 23 24 25 --rotate 1 --warm 3 > gpurun_out/trace_rot_burst.txt 2>&1; tail -12 gpurun_out/trace_rot_burst.txt
python tools/block_trace.py --n 28 --kpos 20 21 22 23 24 25 --rotate 1 --warm 1500 > gpurun_out/trace_rot_capped.txt 2>&1; tail -12 gpurun_out/trace_rot_capped.txt
TQB200_BLOCK_WRITEBACK=rotate timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_block6_kernel -s 34 -c 3 -f -o gpurun_out/r02_v3_block6 python tools/block_times.py --n 28 --layers 16 > gpurun_out/ncu_block.log 2>&1; echo "ncu block rc=$?"
