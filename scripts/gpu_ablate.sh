# ablation of the bench workload (experiment only): which part of the kernel costs what
mkdir -p gpurun_out
for d in "" rot photo rot,photo blur eq noise allrot allrot,photo; do
  echo "DROP=$d"
  B200AUG_BENCH_DROP=$d python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_us'], d['value'])"
done 2>&1 | tee gpurun_out/ablate.log
