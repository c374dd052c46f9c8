#!/bin/bash
R=${1:-1250000}; Q=${2:-8192}
for H in 0 1; do
OFX_SEARCH_HIST=$H OFX_LIB_PATH=outfitx_b200/libofx_debug.so OFX_TC_PROF=1 python tools/prof_step.py --skip-cp --rows $R --queries $Q 2>&1 | grep -E "search prof|  unit" | tail -18 | head -7
done
run() { echo "== $*"; env "$@" python tools/prof_step.py --skip-cp --rows $R --queries $Q --reps 5 2>&1 | grep -E "search ms|rror"; }
run OFX_SEARCH_HIST=1
run OFX_SEARCH_HIST=0
run OFX_SEARCH_HIST=1
run OFX_SEARCH_HIST=0
