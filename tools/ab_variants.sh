python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in default swz0; do
  echo "== $v"; if [ $v = swz0 ]; then export TQB200_LIB=$PWD/torchquantum_b200/lib/exp_swz0.so; fi
  python tools/cfg2_kernels.py 2>&1 | grep cfg2
  python tools/pass_sweep.py --n 28 --mix rot --ks 1 2 4 8 24 64 2>&1 | tail -6
  python tools/pass_sweep.py --n 26 --bwd --mix rot --ks 1 8 24 2>&1 | tail -3
done
