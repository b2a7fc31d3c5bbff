#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels added in round 2's last sessions
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/umma_probe.py --n 16 --time-n 0 --out gpurun_out/x.json > gpurun_out/sanitizer_block_rot.log 2>&1; echo "memcheck block rc=$?"; tail -3 gpurun_out/sanitizer_block_rot.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "quadphase or collapse or write_back" > gpurun_out/sanitizer_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/sanitizer_tests.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "quadphase_kernel" > gpurun_out/racecheck_quadphase.log 2>&1; echo "racecheck quadphase rc=$?"; tail -4 gpurun_out/racecheck_quadphase.log
