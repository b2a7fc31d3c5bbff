#!/bin/bash
# A/B of the forward-only op forms (scaled rotations + folded CNOTs) on the n=28 pass sweep, plus parity and bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== fast forms ON"  > gpurun_out/pass_sweep_ab.log
python tools/pass_sweep.py --n 28 >> gpurun_out/pass_sweep_ab.log 2>&1
echo "== fast forms OFF" >> gpurun_out/pass_sweep_ab.log
TQB200_FAST_FWD=0 python tools/pass_sweep.py --n 28 >> gpurun_out/pass_sweep_ab.log 2>&1
cat gpurun_out/pass_sweep_ab.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['eager'])
r = d['roofline']; print(r['achieved'], r['frac'], r['avg_launch_ms'], r['workload'], r['gate_amp_updates_per_s']); print(r['single_pass_sweep'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
