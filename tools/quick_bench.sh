#!/bin/bash
# usage: tools/quick_bench.sh <tag> [lib.so]  -> prints kernel ms per step for C2
if [ -n "$2" ]; then export ARC_RAD_LIB=$PWD/$2; fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/qb_$1.json 2> gpurun_out/qb_$1.err || tail -5 gpurun_out/qb_$1.err
python -c "
import json,sys; d=json.load(open('gpurun_out/qb_$1.json')); print('$1', round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
