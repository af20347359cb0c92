#!/bin/bash
# GPU session: parity tests, then bench at 1, 2 and 4 launch chains per GPU (run through gpurun from the repo root)
R=${1:-r02b}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_$R.log
for G in 1 2 4; do
  python bench.py --groups $G --no-cpu-baseline > $O/bench_g${G}_$R.json 2> $O/bench_g${G}_$R.err; echo "bench groups=$G rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_g${G}_$R.json"))
    print("groups $G: value %.0f ms/step %.1f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k:round(v["ms_total"],1) for k,v in d["kernels"].items()})
except Exception as e: print("no json", e)
PY
done
