#!/bin/bash
# Standard GPU check used during development: parity tests, a short bench line and the per-CTA trace.
# usage (through gpurun): bash scripts/gpu_check.sh <tag> [pytest-args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench.json"))
    print("value %.3fM samples/s  step %.1f us  frac %.3f  e2e %.0f" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
timeout 120 python scripts/trace_ctas.py > gpurun_out/${tag}_trace.log 2>&1; echo "trace rc=$?"; cat gpurun_out/${tag}_trace.log | tail -25
