#!/bin/bash
# needs the A/B build: `make experiments`, then copy scratch/libacino_b200_experiments.so over acinoset_b200/libacino_b200.so
# for the run (the product library has no environment switches)
# throughput vs resident CTAs per SM of fte_eval (extra dynamic shared memory per CTA lowers the residency)
for padk in "$@"; do
  ACINO_FTE_SMEM_PAD=$((padk*1024)) python bench.py --steps 10 --no-cpu-baseline --no-lm 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('smem pad ${padk} KB', '%.4g frames/s' % d['value'], '%.4f ms' % d['ms_per_step'])"
done
