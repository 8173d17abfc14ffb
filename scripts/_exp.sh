timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/host_profile.py 2>&1 | head -22
timeout 300 python scripts/train_bench.py --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('samples_per_s','ms_per_step','ms_train_only','ms_aug_from_pinned_host','overhead_of_aug_when_overlapped_ms')})"
