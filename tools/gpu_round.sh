#!/bin/bash
# One GPU box visit: parity tests, bench, launch list of one graph replay, optional extras.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --profile-from-start off --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_step.csv python tools/profile_graph_step.py > gpurun_out/launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_graph_step.csv 1 > gpurun_out/launches_graph_step_summary.txt; head -24 gpurun_out/launches_graph_step_summary.txt
for extra in "$@"; do
  case $extra in
    configs) timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; tail -12 gpurun_out/configs.log;;
    ncu) timeout 400 ncu --profile-from-start off --graph-profiling node --set full --clock-control none --import-source on -k regex:'grid_nn_kernel|grid_build_kernel|pointnet_phase_kernel|linear_bf16_kernel|attention_kernel|pose_head_kernel' -o gpurun_out/top_full -f python tools/profile_graph_step.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log;;
    occ) bash tools/sweep_occ.sh "1.0 1.4" "1 1.5 2 3 4 6" > gpurun_out/sweep_occ.txt 2>&1; cat gpurun_out/sweep_occ.txt;;
    train) timeout 600 python tools/bench_train_step.py > gpurun_out/train_step.log 2>&1; tail -4 gpurun_out/train_step.log;;
    trainprof) timeout 600 python tools/profile_train_step.py > gpurun_out/train_profile.txt 2>&1; head -45 gpurun_out/train_profile.txt | cut -c1-200;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log;;
    dgcnn) timeout 600 python tools/run_dgl_dgcnn.py 32 3 > gpurun_out/dgl_dgcnn.txt 2>&1; grep -v Warn gpurun_out/dgl_dgcnn.txt | head -40;;
    pndebug) MPA_PN_DEBUG=1 timeout 300 python tools/pn_debug.py > gpurun_out/pn_debug.txt 2>&1; grep "pn phase 5" gpurun_out/pn_debug.txt | tail -9;;
    refbench) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json;;
  esac
done
