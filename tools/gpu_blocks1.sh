# first GPU check of the tensor-core block path: probe (numerics, norm drift, timing), depth parity, config 4
set -x
timeout 300 python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/umma_probe.json > gpurun_out/umma_probe.log 2>&1; tail -12 gpurun_out/umma_probe.log
timeout 900 python tools/depth_parity.py --n 20 22 24 --depth 100 --full 28 --out gpurun_out/depth_parity.json > gpurun_out/depth_parity.log 2>&1; tail -15 gpurun_out/depth_parity.log
timeout 600 python tools/configs.py --cfg 4 --out gpurun_out/configs_cfg4.json > gpurun_out/configs_cfg4.log 2>&1; tail -5 gpurun_out/configs_cfg4.log
