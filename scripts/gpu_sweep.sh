# launch-geometry sweep of the bench workload (experiment only)
mkdir -p gpurun_out
python neuralnet-tracker-traincode_b200/build.py > /dev/null
for cl in 1 2 4; do for rb in 0 2544; do
  echo -n "cluster=$cl rowbuf=$rb: "
  B200AUG_BENCH_CLUSTER=$cl python bench.py --steps 100 --warmup 5 --no-cpu-baseline --rowbuf $rb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_us'])"
done; done 2>&1 | tee gpurun_out/sweep.log
