#!/bin/bash
# needs the A/B build: `make experiments`, then copy scratch/libacino_b200_experiments.so over acinoset_b200/libacino_b200.so
# for the run (the product library has no environment switches)
# A/B of the fte_eval experiments (ACINO_FTE_EXP bit 0: L2 prefetch of a later tile, bit 1: MUFU sin/cos)
for v in "$@"; do
  ACINO_FTE_EXP=$v python bench.py --steps 20 --no-cpu-baseline --no-lm 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp $v', '%.4g frames/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'single-seq %.1f us' % d['config']['single_sequence_1000f_us_per_launch'], 'e2e %.4g' % d['e2e']['value'])"
done
