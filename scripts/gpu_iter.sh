# one development iteration on the GPU box: parity tests, per-CTA trace, short bench line; $1 = tag
mkdir -p gpurun_out
python neuralnet-tracker-traincode_b200/build.py > /dev/null
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/$1_pytest.log
timeout 120 python scripts/trace_ctas.py 2>&1 | head -10 | tee gpurun_out/$1_trace.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/$1_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_us', d['roofline']['kernel_us'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'whole', d['e2e']['whole_frames']['value'], 'sync', d['e2e']['whole_frames_synchronised_every_step']['value'], 'zero-copy', d['e2e']['zero_copy_frames']['value'])"
