#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python tools/configs.py --cfg 1 3 4 --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; tail -4 gpurun_out/configs.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_block6 -s 12 -c 2 -f -o gpurun_out/r02_v2_block6 python tools/umma_probe.py --n 0 --time-n 28 > gpurun_out/ncu_block6.log 2>&1; echo "ncu block rc=$?"; tail -3 gpurun_out/ncu_block6.log
timeout 900 python -m pytest tests -m gpu -q -x -k "host_mirror" > gpurun_out/pytest_gpu_api.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_api.log
