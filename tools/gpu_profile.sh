# Evidence pass on one B200: launch list of the benchmark step + ncu --set full of the simulator and NAF kernels.
# Usage: bash tools/gpu_profile.sh <tag>    (numbers printed under a profiler are never bench values)
tag=${1:-rX}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-graph > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_dynamics|sim_minv|sim_solve" -s 9 -c 3 -o gpurun_out/${tag}_sim4096_full -f python tools/prof_sim.py 4096 6 > gpurun_out/${tag}_ncu_sim4096.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_dynamics|sim_minv|sim_solve" -s 9 -c 3 -o gpurun_out/${tag}_sim131072_full -f python tools/prof_sim.py 131072 6 > gpurun_out/${tag}_ncu_sim131072.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"policy_act_tc|trunk_tc_layer2|naf_head_kernel|splitk_adam|bn_bwd_apply_dw1|bn_bwd_reduce|replay_append" -s 60 -c 9 -o gpurun_out/${tag}_naf_full -f python bench.py --steps 4 --warmup 6 --no-cpu --no-graph > gpurun_out/${tag}_ncu_naf.log 2>&1
tail -2 gpurun_out/${tag}_ncu_naf.log
