#!/bin/bash
# needs the A/B build: `make experiments`, then copy scratch/libacino_b200_experiments.so over acinoset_b200/libacino_b200.so
# for the run (the product library has no environment switches)
# A/B of the host-pipeline chunk size of acino_fte_eval (frames per chunk): prints e2e frames/s
for c in "$@"; do
  ACINO_E2E_CHUNK=$c python bench.py --steps 5 --no-cpu-baseline --no-lm 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c', 'e2e %.4g frames/s' % d['e2e']['value'], 'value %.4g' % d['value'])"
done
