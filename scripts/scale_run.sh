#!/bin/bash
# Scaling record for profiles/: the contract bench (augmentation), the training step (configs 3/4) and the localizer shape
# (config 5) at N GPUs.  usage (through gpurun --gpus N): bash scripts/scale_run.sh <tag> <N>
tag=$1; n=$2
mkdir -p gpurun_out
run() { if [ "$n" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 "$@"; fi; }
run bench.py --gpus $n --steps 20 --warmup 5 --quick 2>gpurun_out/${tag}_n${n}_bench.err | tail -1 > gpurun_out/${tag}_n${n}_bench.json
run bench.py --gpus $n --workload train 2>gpurun_out/${tag}_n${n}_train.err | tail -1 > gpurun_out/${tag}_n${n}_train.json
run scripts/localizer_bench.py 2>gpurun_out/${tag}_n${n}_loc.err | tail -1 > gpurun_out/${tag}_n${n}_localizer.json
nvidia-smi topo -m > gpurun_out/${tag}_n${n}_topo.txt 2>&1; nproc >> gpurun_out/${tag}_n${n}_topo.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/${tag}_n${n}_topo.txt
python - <<PY
import json
for k in ("bench", "train", "localizer"):
    try:
        d = json.load(open("gpurun_out/${tag}_n${n}_%s.json" % k))
        if k == "bench":
            print(k, "value %.3fM e2e %.0f per-rank ms %s h2d %s host_ms %s" % (d["value"]/1e6, d["e2e"]["value"], [round(x, 4) for x in d["details"]["per_rank_ms_per_step"]],
                  [round(x, 1) for x in d["e2e"]["per_rank"]["h2d_GBps"]], [round(x, 2) for x in d["e2e"]["per_rank"]["host_ms_per_step"]]))
        else:
            print(k, {a: d[a] for a in d if a in ("samples_per_s", "ms_per_step", "ms_train_only", "kernel_us", "frac", "overhead_of_aug_when_overlapped_ms")})
    except Exception as e:
        print(k, "failed", e)
PY
