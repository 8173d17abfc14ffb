# evidence pass for profiles/: $1 = tag (e.g. r01d).  GPU tests, smoke, bench line (with cpu_baseline), reference arm, ncu launch list.
mkdir -p gpurun_out
python neuralnet-tracker-traincode_b200/build.py > /dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1; nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/$1_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/$1_smoke.log
python bench.py --steps 200 --warmup 10 2>gpurun_out/$1_bench.err | tail -1 > gpurun_out/$1_bench.json; cut -c1-300 gpurun_out/$1_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/$1_bench_reference.json; cut -c1-200 gpurun_out/$1_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 8 --warmup 3 --spin-s 0 --no-cpu-baseline > gpurun_out/$1_ncu_bench.log 2>&1
tail -3 gpurun_out/$1_launches.csv | cut -c1-200
