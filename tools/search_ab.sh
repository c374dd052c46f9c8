#!/bin/bash
# A/B of the counting bound on the shipped search: bash tools/search_ab.sh ROWS [QUERIES]
R=${1:-1250000}; Q=${2:-8192}
run() { echo "== $*"; env "$@" python tools/prof_step.py --skip-cp --rows $R --queries $Q --reps 5 2>&1 | grep -E "search ms|rror"; }
run OFX_SEARCH_HIST=0
run OFX_SEARCH_HIST=1
run OFX_SEARCH_HIST=0
run OFX_SEARCH_HIST=1
OFX_LIB_PATH=outfitx_b200/libofx_debug.so OFX_TC_PROF=1 python tools/prof_step.py --skip-cp --rows $R --queries $Q 2>&1 | grep -E "search prof" | grep "rows=$R" | tail -1
OFX_LIB_PATH=outfitx_b200/libofx_debug.so OFX_MERGE_PROF=1 python tools/prof_step.py --skip-cp --rows $R --queries $Q 2>&1 | grep -E "merge prof" | tail -1
