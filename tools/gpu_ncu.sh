#!/bin/bash
# ncu --set full captures of the gate-pass kernels: cfg2 (on-chip, fwd + adjoint bwd) and the n=28 streamed passes.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -c 2 -f -o gpurun_out/cfg2_pass python tools/perf_probe.py --big > gpurun_out/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -c 7 -f -o gpurun_out/n28_pass python tools/perf_probe.py --skip-cfg2 --big 28 > gpurun_out/ncu_n28.log 2>&1; echo "ncu n28 rc=$?"
python tools/perf_probe.py --big 24 26 28 > gpurun_out/perf_probe.log 2>&1; tail -5 gpurun_out/perf_probe.log
python tools/pass_sweep.py --n 28 > gpurun_out/pass_sweep.log 2>&1; tail -12 gpurun_out/pass_sweep.log
