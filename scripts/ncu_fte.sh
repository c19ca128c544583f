#!/bin/bash
# one full ncu capture of a 256 000-frame fte_eval launch (+ source page) -> gpurun_out/$1.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:fte_eval -s ${2:-4} -c 1 -f -o gpurun_out/$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lm --no-sba > gpurun_out/$1.log 2>&1
tail -3 gpurun_out/$1.log
