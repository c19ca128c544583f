#!/bin/bash
# A/B of fte_eval builds with the parity tests in front: scripts/ab_fte_checked.sh lib1.so lib2.so ...
for lib in "$@"; do
  cp "$lib" acinoset_b200/libacino_b200.so
  python -m pytest tests/test_fte_eval_gpu.py -x -q -m gpu 2>&1 | tail -1
done
for rep in 1 2; do scripts/ab_fte.sh "$@"; done
