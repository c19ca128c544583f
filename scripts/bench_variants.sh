#!/bin/bash
# needs the A/B build: `make experiments`, then copy scratch/libacino_b200_experiments.so over acinoset_b200/libacino_b200.so
# for the run (the product library has no environment switches)
# quick A/B of fte_eval kernel variants (ACINO_FTE_VARIANT) - prints frames/s and ms/step
for v in "$@"; do
  ACINO_FTE_VARIANT=$v python bench.py --steps 10 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v', '%.4g frames/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'single-seq %.1f us' % d['config']['single_sequence_1000f_us_per_launch'])"
done
