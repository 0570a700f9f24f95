#!/bin/bash
# usage: tools/quick_bench.sh <tag> [lib.so] [extra bench args...]  -> prints kernel ms per step
tag=$1; shift
if [ -n "$1" ] && [ -f "$1" ]; then export ARC_RAD_LIB=$PWD/$1; shift; fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-aer --no-extras "$@" > gpurun_out/qb_$tag.json 2> gpurun_out/qb_$tag.err || tail -5 gpurun_out/qb_$tag.err
python -c "
import json,sys; d=json.load(open('gpurun_out/qb_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
