#!/bin/bash
# Experiment (profiles/r01_grid_sweep.md): grid_nn_kernel at 3 / 5 / 6 resident CTAs per SM.
# Build the variants first, in multi_part_assembly_b200/csrc:
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DMPA_NN_MIN_CTAS=$v -c chamfer.cu -o build/chamfer_v$v.o
#   nvcc -shared -o libmpa_exp$v.so build/{mpa_runtime,chamfer_v$v,se3,pointnet,pointnet_bwd,linear,knn,loss}.o
# (MPA_B200_LIB points the Python binding at an alternative library.)
for v in 3 5 6; do
  MPA_B200_LIB=$PWD/multi_part_assembly_b200/csrc/libmpa_exp$v.so python bench.py --steps 100 --warmup 5 --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']
print('min_ctas=$v step %.4f ms | nn_shape %.3f nn_part %.3f' % (d['ms_per_step'], k['chamfer_grid_nn_shape'], k['chamfer_grid_nn_part']))"
done
python bench.py --steps 100 --warmup 5 --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']
print('min_ctas=4 (default) step %.4f ms | nn_shape %.3f nn_part %.3f' % (d['ms_per_step'], k['chamfer_grid_nn_shape'], k['chamfer_grid_nn_part']))"
