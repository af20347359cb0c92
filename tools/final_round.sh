#!/bin/bash
# Last GPU session of the round: tests, the record bench lines, the CPU baseline in full, ncu of the assembly kernel.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -s > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$R.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$R.log 2>&1; cat $O/smoke_$R.log
python bench.py > $O/bench_final_$R.json 2> $O/bench_final_$R.err; echo "bench rc=$?"
python bench.py --nfreq 2 --no-cpu-baseline > $O/bench_f2_$R.json 2> $O/bench_f2_$R.err
python bench.py --config cfg2 --no-cpu-baseline > $O/bench_cfg2_$R.json 2> $O/bench_cfg2_$R.err
python bench.py --config cfg4 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_cfg4_$R.json 2> $O/bench_cfg4_$R.err
python bench.py --config cfg4 --dtype c128 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_cfg4_c128_$R.json 2> $O/bench_cfg4_c128_$R.err
python tools/cpu_baseline.py > $O/cpu_baseline_$R.json 2> $O/cpu_baseline_$R.err
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --groups 1"
ncu --set full --clock-control none -k regex:assemble_kernel -c 1 -o $O/prof_assemble_$R -f $FULL > $O/ncu_assemble_$R.log 2>&1
python - <<PY
import json
for t in ("bench_final","bench_f2","bench_cfg2","bench_cfg4","bench_cfg4_c128"):
    try:
        d=json.loads([l for l in open("$O/%s_$R.json"%t).read().splitlines() if l.startswith("{")][-1])
        print(t, "value %.0f ms %.1f e2e %.0f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("checks"), {k:(round(v["ms_total"],1), round(v.get("frac",0),3)) for k,v in d["kernels"].items() if v.get("ms_total",0)>0.05})
    except Exception as e: print(t, "ERR", e)
print(open("$O/cpu_baseline_$R.json").read()[:1500])
PY
