#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (run through gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/sanitize_round.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
K='(test_solve_forward_adjoint and c64) or (test_fwi_loss_and_grad and c64 and 48)'
for TOOL in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $TOOL --print-limit 30 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/sanitizer_${TOOL}_$R.log 2>&1
  echo "$TOOL exit $?" >> $O/sanitizer_${TOOL}_$R.log
done
tail -n 8 $O/sanitizer_*_$R.log
