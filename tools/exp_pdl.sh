for nf in 2 4 8 16; do
for v in 0 4 8 16; do
  echo -n "nfreq $nf pdl_min $v: "
  UST_PDL_MIN_BATCH=$v python bench.py --nfreq $nf --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.1f' % (d['ms_per_step']))"
done; done
