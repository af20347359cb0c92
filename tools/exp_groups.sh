# experiment: programmatic dependent launch on / off at several batch sizes (run through gpurun from the repo root)
run() { echo "== $*"; env "${@:2}" python bench.py $1 --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ms/step %.1f' % (d['ms_per_step']))"; }
run "--nfreq 16" UST_NO_PDL=0
run "--nfreq 16" UST_NO_PDL=1
run "--nfreq 8" UST_NO_PDL=0
run "--nfreq 8" UST_NO_PDL=1
run "--nfreq 4" UST_NO_PDL=0
run "--nfreq 4" UST_NO_PDL=1
run "--config cfg2" UST_NO_PDL=0
run "--config cfg2" UST_NO_PDL=1
run "--config cfg4" UST_NO_PDL=0
run "--config cfg4" UST_NO_PDL=1
