#!/bin/bash
# round 2, GPU visit A: parity suite with the tightened bars, block-kernel ncu capture, cfg 3/4 at full size, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/configs.py --cfg 3 4 --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; tail -4 gpurun_out/configs.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_cfg4_launches.csv python tools/configs.py --cfg 4 > gpurun_out/cfg4_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_block6 -s 40 -c 3 -f -o gpurun_out/r02_block6 python tools/umma_probe.py --n 0 --time-n 28 > gpurun_out/ncu_block6.log 2>&1; echo "ncu block rc=$?"; tail -3 gpurun_out/ncu_block6.log
