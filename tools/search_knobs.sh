#!/bin/bash
# Scheduler knob sweep of the shipped search on one shard size: bash tools/search_knobs.sh ROWS [QUERIES]
R=${1:-1250000}; Q=${2:-8192}
run() { echo "== $*"; env "$@" python tools/prof_step.py --skip-cp --rows $R --queries $Q --reps 5 2>&1 | grep "search ms" | cut -c1-60; }
run OFX_SEARCH_WINDOW=8
run OFX_SEARCH_WINDOW=4
run OFX_SEARCH_WINDOW=12
run OFX_SEARCH_WINDOW=16
run OFX_SEARCH_WINDOW=8 OFX_SEARCH_CHECK=4
run OFX_SEARCH_WINDOW=16 OFX_SEARCH_CHECK=16
run OFX_SEARCH_PREFETCH=4
run OFX_SEARCH_PREFETCH=12
run OFX_SEARCH_PREFETCH=16 OFX_SEARCH_WINDOW=16
run OFX_SEARCH_LEAD=0
run OFX_SEARCH_WINDOW=8
