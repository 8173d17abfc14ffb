mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1_pytest_gpu.log; tail -15 gpurun_out/r1_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r1_smoke.log
python bench.py --steps 200 --warmup 10 2>gpurun_out/r1_bench.err | tee gpurun_out/r1_bench.json
tail -5 gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r1_bench_ref.err | tee gpurun_out/r1_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 8 --warmup 3 --spin-s 0 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_augment -s 3 -c 2 -o gpurun_out/r1_fused python bench.py --steps 4 --warmup 3 --spin-s 0 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
