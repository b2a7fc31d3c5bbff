#!/bin/bash
# 2-GPU visit: sharded-statevector parity (n=20, 24), cfg-5 family at n=30/31 on 2 ranks, batch-sharded bench at N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_sharded.py --check 20 > gpurun_out/dist_check20.log 2>&1; echo "check20 rc=$?"; tail -2 gpurun_out/dist_check20.log
timeout 600 $TR tools/dist_sharded.py --check 24 > gpurun_out/dist_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist_check24.log
timeout 900 $TR tools/dist_sharded.py --bench 30 --out gpurun_out/cfg5_n30_w2.json > gpurun_out/dist_bench30.log 2>&1; echo "bench30 rc=$?"; tail -1 gpurun_out/dist_bench30.log
timeout 900 $TR tools/dist_sharded.py --bench 32 --out gpurun_out/cfg5_n32_w2.json > gpurun_out/dist_bench32.log 2>&1; echo "bench32 rc=$?"; tail -1 gpurun_out/dist_bench32.log
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; cut -c1-400 gpurun_out/bench_n2.json
