#!/bin/bash
# 8-GPU visit: sharded-statevector parity at g=3 (n=24 vs one GPU) and cfg 5 proper: n=34 over 8 B200.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 300 $TR tools/dist_sharded.py --check 24 > gpurun_out/dist8_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist8_check24.log
timeout 600 $TR tools/dist_sharded.py --bench 34 --out gpurun_out/cfg5_n34_w8.json > gpurun_out/dist8_bench34.log 2>&1; echo "bench34 rc=$?"; tail -1 gpurun_out/dist8_bench34.log
