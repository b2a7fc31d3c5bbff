#!/bin/bash
# ncu evidence for the round-2 final kernels: --set full of the rotating block pass and the quadratic-phase pass (QFT n = 28
# plan), the GEMM (cfg 2 step), and the launch list of the default bench command
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tqb_block6_kernel|tqb_quadphase_kernel" -s 3 -c 6 -f -o gpurun_out/r02_v3_block6_quad python tools/qft_probe2.py 28 > gpurun_out/ncu_qft.log 2>&1; echo "ncu qft rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_v3_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_v3_bench_launches.csv
