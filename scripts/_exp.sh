timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/gpu_stress.py 8 2>&1 | tail -2
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --quick 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step %.1f us' % (d['ms_per_step']*1e3))"
timeout 120 python scripts/trace_ctas.py 2>&1 | grep -E "kernel span|^all|^plain|per-SM"
timeout 300 python scripts/localizer_bench.py 2>&1 | tail -1 | cut -c1-60,200-330
