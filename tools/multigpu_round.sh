#!/bin/bash
# Multi-GPU session (run through gpurun --gpus N from the repo root): bash tools/multigpu_round.sh N r02 [cfg4|nocfg4]
N=${1:-2}
R=${2:-r02}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N > $O/bench_${N}gpu_strong_$R.json 2> $O/bench_${N}gpu_strong_$R.err; echo "strong rc=$?"
if [ "${3:-cfg4}" = "cfg4" ]; then $TR bench.py --gpus $N --config cfg4 --steps 2 --warmup 1 > $O/bench_${N}gpu_cfg4_$R.json 2> $O/bench_${N}gpu_cfg4_$R.err; echo "cfg4 rc=$?"; fi
python - <<PY
import json
for t in ("strong","cfg4"):
    try:
        d=json.loads([l for l in open("$O/bench_${N}gpu_%s_$R.json"%t).read().splitlines() if l.startswith("{")][-1])
        print(t, "N=$N value %.0f ms %.1f e2e %.0f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "weak:", d.get("weak_scaling"))
    except Exception as e: print(t, "ERR", e)
PY
