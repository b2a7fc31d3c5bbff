#!/bin/bash
# Round-end GPU visit: parity suite, smoke, bench (ours + reference arm), ncu launch list of the bench command,
# ncu --set full captures of the dominant kernels (n=28 streamed forward pass, config-2 on-chip forward and adjoint pass).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python tools/configs.py --cfg 1 3 4 --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -s 7 -c 2 -f -o gpurun_out/n28_pass python tools/perf_probe.py --skip-cfg2 --big 28 > gpurun_out/ncu_n28.log 2>&1; echo "ncu n28 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -s 4 -c 2 -f -o gpurun_out/cfg2_pass python tools/cfg2_kernels.py --iters 2 > gpurun_out/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_dense_tc_kernel -s 10 -c 1 -f -o gpurun_out/dense_tc python tools/dense_probe.py --n 26 > gpurun_out/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
python tools/dense_probe.py --n 28 > gpurun_out/dense_probe_n28.log 2>&1
cat gpurun_out/bench.json | cut -c1-1500
