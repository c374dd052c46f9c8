#!/bin/bash
# Scheduler knob sweep of the shipped search on one shard size: bash tools/search_knobs.sh ROWS
R=${1:-1250000}
run() { echo "== $*"; env "$@" python tools/prof_step.py --skip-cp --rows $R --reps 5 2>&1 | grep "search ms"; }
run OFX_SEARCH_LEAD=0
run OFX_SEARCH_LEAD=1
run OFX_SEARCH_LEAD=3
run OFX_SEARCH_LEAD=3 OFX_SEARCH_WINDOW=0
run OFX_SEARCH_LEAD=3 OFX_SEARCH_WINDOW=16
run OFX_SEARCH_LEAD=3 OFX_SEARCH_WINDOW=32
run OFX_SEARCH_LEAD=3 OFX_SEARCH_WINDOW=16 OFX_SEARCH_CHECK=4
run OFX_SEARCH_LEAD=3 OFX_SEARCH_PREFETCH=16 OFX_SEARCH_WINDOW=16
run OFX_SEARCH_LEAD=3 OFX_SEARCH_PREFETCH=4
run OFX_SEARCH_PREFETCH=0
echo "== debug counters"
OFX_LIB_PATH=outfitx_b200/libofx_debug.so OFX_TC_PROF=1 python tools/prof_step.py --skip-cp --rows $R 2>&1 | grep -E "search prof|search ms" | tail -4
OFX_LIB_PATH=outfitx_b200/libofx_debug.so OFX_TC_PROF=1 OFX_SEARCH_LEAD=0 python tools/prof_step.py --skip-cp --rows $R 2>&1 | grep -E "search prof|search ms" | tail -4
