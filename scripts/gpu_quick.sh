# quick GPU check: parity tests + smoke + a short bench line (no CPU baseline)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/quick_pytest.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/quick_bench.json
