echo "== default (bwd minb 2, R_GRAD 4)"; python tools/perf_probe.py --big 2>&1 | tail -1
echo "== R_GRAD 3"; TQB200_RGRAD=3 python tools/perf_probe.py --big 2>&1 | tail -1
echo "== bwd minb 1, R_GRAD 4"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_bwd1.so python tools/perf_probe.py --big 2>&1 | tail -1
echo "== bwd sweep default"; python tools/pass_sweep.py --n 26 --bwd --ks 1 8 24 --mix rot 2>&1 | tail -3
echo "== bwd sweep R_GRAD 3"; TQB200_RGRAD=3 python tools/pass_sweep.py --n 26 --bwd --ks 1 8 24 --mix rot 2>&1 | tail -3
echo "== bwd sweep minb 1"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_bwd1.so python tools/pass_sweep.py --n 26 --bwd --ks 1 8 24 --mix rot 2>&1 | tail -3
