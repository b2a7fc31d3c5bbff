#!/bin/bash
# ncu --set full captures of the gate-pass kernels: n=28 streamed forward passes and the cfg2 on-chip fwd/bwd kernels.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -s 7 -c 2 -f -o gpurun_out/n28_pass python tools/perf_probe.py --skip-cfg2 --big 28 > gpurun_out/ncu_n28.log 2>&1; echo "ncu n28 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -s 4 -c 2 -f -o gpurun_out/cfg2_pass python tools/perf_probe.py --big > gpurun_out/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
ls -la gpurun_out/*.ncu-rep
