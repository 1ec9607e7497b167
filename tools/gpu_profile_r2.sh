# Round-2 evidence pass on one B200 (numbers printed under a profiler are never bench values).
# Usage: bash tools/gpu_profile_r2.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err
RLOA_CONTACTS=0 python bench.py --steps 200 --warmup 20 --no-cpu --no-extras > gpurun_out/${tag}_bench_n1_free_dynamics.json 2>> gpurun_out/${tag}_bench_n1.err
python bench.py --trunk fp32 --steps 200 --warmup 20 --no-cpu --no-extras > gpurun_out/${tag}_bench_n1_fp32.json 2>> gpurun_out/${tag}_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-graph --no-extras --preroll 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"naf_learn_cluster|learn_pack" -s 8 -c 2 -o gpurun_out/${tag}_learn_cluster_full -f python tools/learn_timing.py > gpurun_out/${tag}_ncu_learn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_dynamics|sim_minv|sim_solve" -s 9 -c 3 -o gpurun_out/${tag}_sim4096_full -f python tools/prof_sim.py 4096 6 > gpurun_out/${tag}_ncu_sim4096.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sim_solve -s 395 -c 1 -o gpurun_out/${tag}_solve_contacts_full -f python tools/prof_contacts.py 350 > gpurun_out/${tag}_ncu_solve_contacts.log 2>&1
for f in learn_cluster_full sim4096_full solve_contacts_full; do ncu -i gpurun_out/${tag}_${f}.ncu-rep --page raw --csv > gpurun_out/${tag}_${f}_raw.csv 2>/dev/null; done
python tools/prof_contacts.py 900 graph 2>&1 | grep arms > gpurun_out/${tag}_contacts_over_episodes.txt
python tools/learn_cluster_profile.py 1024 > gpurun_out/${tag}_learn_timeline_warm.txt 2>&1
python tools/learn_cluster_profile.py 1024 flush > gpurun_out/${tag}_learn_timeline_flushed.txt 2>&1
python tools/learn_timing.py 1024 > gpurun_out/${tag}_learn_timing.txt 2>&1
RLOA_LEARN_CLUSTER=0 python tools/learn_timing.py 1024 > gpurun_out/${tag}_learn_timing_multilaunch.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${tag}_memcheck.log 2>&1; tail -2 gpurun_out/${tag}_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${tag}_racecheck.log 2>&1; tail -2 gpurun_out/${tag}_racecheck.log
tail -c 400 gpurun_out/${tag}_bench_n1.json; cat gpurun_out/${tag}_bench_reference.json | cut -c1-600
