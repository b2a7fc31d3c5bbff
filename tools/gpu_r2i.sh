#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for m in 4 8; do
  echo "== TQB200_BLOCK_SPLIT=$m"
  TQB200_BLOCK_SPLIT=$m timeout 300 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe_s$m.json > gpurun_out/umma_probe_s$m.log 2>&1; echo "probe rc=$?"; grep -v "^\[" gpurun_out/umma_probe_s$m.log | tail -7; grep -c "ERROR\|differs" gpurun_out/umma_probe_s$m.log; sort -t' ' -k8 -g gpurun_out/umma_probe_s$m.log | grep max_rel | tail -1
done
TQB200_BLOCK_SPLIT=8 python tools/block_trace.py 2>&1 | tail -11
timeout 600 $TR tools/dist_sharded.py --check 24 --exchange p2p > gpurun_out/dist_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist_check24.log
timeout 900 $TR tools/dist_sharded.py --bench 32 --exchange p2p --out gpurun_out/r02_cfg5_n32_w2.json > gpurun_out/dist_bench32.log 2>&1; echo "bench32 rc=$?"; tail -1 gpurun_out/dist_bench32.log | cut -c1-1500
