#!/bin/bash
# sweep of the grid sizing knobs on the pose-Chamfer timing (scratch)
for eff in 0 1; do for occ in 1.5 3 6 12 24; do
  echo "eff=$eff occ=$occ: $(MPA_GRID_EFF=$eff MPA_GRID_OCC=$occ python tools/time_chamfer.py pose 2>&1 | grep pose_chamfer | tr '\n' ' ')"
done; done
