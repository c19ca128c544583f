#!/bin/bash
# A/B of fte_eval builds on the GPU box: scripts/ab_fte.sh lib1.so lib2.so ...  (each is copied over the product library
# for one short bench run; prints frames/s, ms/step, single-sequence latency)
for lib in "$@"; do
  cp "$lib" acinoset_b200/libacino_b200.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-lm --no-sba 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '%.4g frames/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'single-seq %.2f us' % d['config']['single_sequence_1000f_us_per_launch'], 'e2e %.4g' % d['e2e']['value'])"
done
