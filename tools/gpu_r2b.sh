#!/bin/bash
# round 2, GPU visit B: block kernel after the instruction diet (probe: numerics + timing), full parity suite through the
# torch.library ops, bench
mkdir -p gpurun_out
timeout 300 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe.json > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"; grep -v "^\[" gpurun_out/umma_probe.log | tail -12; grep -c "inplace_equal\|ERROR\|differs" gpurun_out/umma_probe.log; sort -t' ' -k8 -g gpurun_out/umma_probe.log | grep max_rel | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python tools/configs.py --cfg 4 --out gpurun_out/configs_cfg4.json > gpurun_out/configs_cfg4.log 2>&1; echo "configs rc=$?"; tail -2 gpurun_out/configs_cfg4.log
