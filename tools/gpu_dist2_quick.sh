mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests -m gpu -x -q -k "qft or forward_only or sharded" 2>&1 | tail -2
timeout 600 $TR tools/dist_sharded.py --check 24 > gpurun_out/dist_check24.log 2>&1; echo "check24 rc=$?"; tail -1 gpurun_out/dist_check24.log
timeout 900 $TR tools/dist_sharded.py --bench 32 --out gpurun_out/cfg5_n32_w2.json > gpurun_out/dist_bench32.log 2>&1; echo "bench32 rc=$?"; tail -1 gpurun_out/dist_bench32.log
