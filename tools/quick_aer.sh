#!/bin/bash
# usage: tools/quick_aer.sh <tag> [lib.so]  -> aerosol optics kernel ms on the 128x64x50 profiling tile and on C2
tag=$1; shift
if [ -n "$1" ] && [ -f "$1" ]; then export ARC_RAD_LIB=$PWD/$1; shift; fi
python tools/prof_aer.py 425 300 50 2>&1 | tail -1 | sed "s/^/$tag /"
