echo "== small-CTA bwd minb 3 (168 regs)"; python tools/perf_probe.py --big 2>&1 | tail -1
echo "== small-CTA bwd minb 2 (255 regs)"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_s2.so python tools/perf_probe.py --big 2>&1 | tail -1
echo "== small-CTA bwd minb 4 (128 regs)"; TQB200_LIB=$PWD/torchquantum_b200/lib/exp_s4.so python tools/perf_probe.py --big 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
