#!/bin/bash
# 2 GPUs: sharded tests (both exchange back ends), cfg 5 family at n = 32, bench.py --gpus 2 (sharded_configs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests -m gpu -x -q -k "sharded" 2>&1 | tail -3
timeout 600 $TR tools/dist_sharded.py --check 24 > gpurun_out/dist_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist_check24.log
timeout 900 $TR tools/dist_sharded.py --bench 32 --out gpurun_out/cfg5_n32_w2.json > gpurun_out/dist_bench32.log 2>&1; echo "bench32 rc=$?"; tail -1 gpurun_out/dist_bench32.log
timeout 900 $TR bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print(d["value"], d["ms_per_step"], json.dumps(d.get("sharded_configs"))[:1500])
PY
