#!/bin/bash
# A/B of the step-level changes on ONE box: bench.py (cfg C, graph replay) with each switch.
run() { echo "== $1"; env $1 python bench.py --steps 300 --warmup 10 --no-extra --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
run "X=1"
run "MPA_NO_FFN_BLOCK=1"
run "MPA_NO_PREPARE=1"
run "MPA_NO_PREFETCH=1"
run "MPA_NO_PREPARE=1 MPA_NO_PREFETCH=1"
run "MPA_NO_PREPARE=1 MPA_NO_PREFETCH=1 MPA_NO_FFN_BLOCK=1"
run "X=1"
