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
