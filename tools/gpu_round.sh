#!/bin/bash
# GPU session: parity tests, then bench variants (run through gpurun from the repo root)
R=${1:-r02d}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -s -x > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu_$R.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.0f ms/step %.1f e2e %.0f dev GB %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["device_bytes"]/1e9), {k:round(v["ms_total"],1) for k,v in d["kernels"].items()}, "grad frac", d["kernels"].get("gradient",{}).get("frac"))
except Exception as e: print("no json", sys.argv[1], e)
PY
}
python bench.py --no-cpu-baseline > $O/bench_$R.json 2> $O/bench_$R.err; summ $O/bench_$R.json
for G in 1 2; do
  python bench.py --nfreq 2 --groups $G --no-cpu-baseline > $O/bench_f2_g${G}_$R.json 2> $O/bench_f2_g${G}_$R.err; summ $O/bench_f2_g${G}_$R.json
done
python bench.py --config cfg2 --no-cpu-baseline > $O/bench_cfg2_$R.json 2> $O/bench_cfg2_$R.err; summ $O/bench_cfg2_$R.json
