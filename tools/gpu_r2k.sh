#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/configs.py --cfg 4 --out gpurun_out/configs_cfg4.json > gpurun_out/configs_cfg4.log 2>&1; echo "cfg4 rc=$?"; tail -1 gpurun_out/configs_cfg4.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eager']['ms_per_step']); r=d['roofline']; print(r['frac'], r['avg_launch_ms'], r['workload'], r['circuit'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_gemm_tf32x3 -s 8 -c 2 -f -o gpurun_out/r02_gemm python tools/gemm_probe.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
