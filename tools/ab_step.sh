#!/bin/bash
# A/B of the step-level switches on ONE box: bench.py (cfg C, graph replay) with each switch.
# The library reads these variables once per process (static), so every run is its own process.
#   MPA_NO_FUSED_ATTN  QKV GEMM + attention kernel instead of encoder_attn_kernel
#   MPA_FFN_CLUSTER=n  cluster width of the FFN block (1 = one CTA per token tile)
#   MPA_PN_NO_STASH    PointNet launches 4, 5 recompute from the points
#   MPA_PN_PHASE1      layer-1 statistics by the MMA launch instead of the point moments
run() { echo "== $1"; env $1 python bench.py --steps 300 --warmup 10 --no-extra --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
run "X=1"
run "MPA_NO_FUSED_ATTN=1"
run "MPA_FFN_CLUSTER=1"
run "MPA_PN_NO_STASH=1"
run "MPA_PN_PHASE1=1"
run "X=1"
