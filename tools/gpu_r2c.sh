#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_dbg.sh 2>&1 | tee gpurun_out/block_ablation.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
