python -m pytest tests/test_sim_gpu.py tests/test_mesh_hull_gpu.py -x -q -m gpu 2>&1 | tail -2
for sp in 1 0; do
RLOA_SIM_SPLIT=$sp timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sim_solve|sim_post" --csv --log-file gpurun_out/r2y_dur_$sp.csv python tools/prof_contacts.py 400 > gpurun_out/r2y_log_$sp.log 2>&1
RLOA_SIM_SPLIT=$sp python bench.py --steps 200 --warmup 20 --no-cpu --no-extras | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split=$sp', d['ms_per_step'], d['phases_ms'], d['warm_l2']['ms_per_step'], d['gpu_launches'])"
done
