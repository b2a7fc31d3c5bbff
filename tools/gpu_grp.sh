for g in 0 1 2 3 5; do echo "== TQB200_TILE_GROUP=$g"; TQB200_TILE_GROUP=$g python tools/umma_probe.py --n 20 --time-n 28 --out gpurun_out/x.json 2>&1 | grep -v "^norm\|^\[" | sed 's/ms_min.*GBps/GBps/'; done
echo "== sweep group 2"; TQB200_TILE_GROUP=2 python tools/kpos_sweep.py | awk 'NR%2==1'
