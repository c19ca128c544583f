#!/bin/bash
# scripts/ab_env.sh lib.so "ENV=.. ENV=.." ...: short bench of one library under several environments
lib=$1; shift
cp "$lib" acinoset_b200/libacino_b200.so
for e in "$@"; do
  env $e python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-lm --no-sba 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$e', '%.4g frames/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'single-seq %.2f us' % d['config']['single_sequence_1000f_us_per_launch'])"
done
