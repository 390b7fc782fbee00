#!/bin/bash
# Chamfer grid sizing sweep on the REAL bench step (untrained pn_transformer, cfg C): step time and
# the four Chamfer kernels per setting.  usage: sweep_occ.sh "<fine values>" "<occ values>"
for fine in $1; do for occ in $2; do
  MPA_GRID_FINE=$fine MPA_GRID_OCC=$occ python bench.py --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']
print('fine=$fine occ=$occ step %.3f ms | nn_shape %.3f build_shape %.3f nn_part %.3f build_part %.3f' % (d['ms_per_step'], k['chamfer_grid_nn_shape'], k['chamfer_grid_build_shape'], k['chamfer_grid_nn_part'], k['chamfer_grid_build_part']))"
done; done
