echo "== default"; python tools/cfg2_kernels.py 2>&1 | tail -4
echo "== small-CTA bwd minb 2 (255 regs)"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_s2.so python tools/cfg2_kernels.py 2>&1 | tail -4
echo "== small-CTA bwd minb 4 (128 regs)"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_s4.so python tools/cfg2_kernels.py 2>&1 | tail -4
python tools/pass_sweep.py --n 28 --mix rot 2>&1 | head -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_pass_kernel -s 4 -c 1 -f -o gpurun_out/cfg2_bwd python tools/cfg2_kernels.py --iters 2 > gpurun_out/ncu_cfg2_bwd.log 2>&1; echo "ncu rc=$?"
