#!/bin/bash
# libtqb200 with the block kernel's ablation switches and pipeline trace compiled in (tools/block_trace.py,
# TQB200_BLOCK_DEBUG): -> tools/_ab/libtqb200_debug.so (git-ignored, travels with gpurun)
cd "$(dirname "$0")/.." && mkdir -p tools/_ab && /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -DTQB_BLOCK_DEBUG \
  -o tools/_ab/libtqb200_debug.so torchquantum_b200/csrc/tqb_cuda.cu
