#!/bin/bash
# scripts/ab_e2e.sh lib.so CHUNK...: end-to-end rate of acino_fte_eval for several host-pipeline chunk sizes (experiments build)
lib=$1; shift
cp "$lib" acinoset_b200/libacino_b200.so
for c in "$@"; do
  ACINO_E2E_CHUNK=$c python bench.py --steps 5 --no-cpu-baseline --no-lm --no-sba 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c', 'e2e %.4g frames/s' % d['e2e']['value'], 'of bus %.3f' % d['e2e']['frac_of_pcie_bound'], 'value %.4g' % d['value'])"
done
