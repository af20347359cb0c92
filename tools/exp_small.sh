# experiment: split-K cluster size of the sweep GEMM at small batches
run() { echo -n "$*: "; env "${@:2}" python bench.py $1 --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.1f  sweep_gemm %.1f ms' % (d['ms_per_step'], d['kernels']['sweep_gemm']['ms_total']))"; }
for a in "--nfreq 2" "--nfreq 4" "--config cfg2"; do
for k in 1 2 4; do run "$a" UST_KSPLIT=$k; done
done
