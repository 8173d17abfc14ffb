for v in "" _v3 _v2; do
  echo "=== lib$v"
  export B200AUG_LIB=$PWD/neuralnet-tracker-traincode_b200/lib/libb200aug$v.so
  for wc in 0 296; do
  B200AUG_WARP_CTAS=$wc timeout 120 python scripts/trace_ctas.py 2>&1 | grep -E "kernel span|canvas workers:|^rotated|per-SM"
  B200AUG_WARP_CTAS=$wc timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --quick 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('warp_ctas $wc step %.1f us' % (d['ms_per_step']*1e3))"
  done
done
