#!/bin/bash
# 8 GPUs: sharded parity at g = 3 (n = 24 vs one GPU), bench.py --gpus 8 (headline weak scaling + cfg 3 strong scaling + cfg 5 n = 34)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/dist_sharded.py --check 24 > gpurun_out/dist8_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist8_check24.log
timeout 900 $TR bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -c 300 gpurun_out/bench_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n8.json"))
print(d["value"], d["ms_per_step"], json.dumps(d.get("sharded_configs"))[:2500])
PY
