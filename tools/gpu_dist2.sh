#!/bin/bash
# 2-GPU visit: sharded-statevector parity over NVLink (p2p stores and NCCL), cfg-5 family at n=32 with both exchange
# backends, bench at N=2 (headline + cfg 3 strong-scaled + cfg 5 sharded).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus" 2>&1 | tail -15
timeout 600 $TR tools/dist_sharded.py --check 24 --exchange p2p > gpurun_out/dist_check24.log 2>&1; echo "check24 rc=$?"; tail -2 gpurun_out/dist_check24.log
timeout 900 $TR tools/dist_sharded.py --bench 32 --exchange both --out gpurun_out/r02_cfg5_n32_w2.json > gpurun_out/dist_bench32.log 2>&1; echo "bench32 rc=$?"; tail -3 gpurun_out/dist_bench32.log | cut -c1-1800
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d.get('sharded_configs'))[:3000])"; tail -3 gpurun_out/bench_n2.err
