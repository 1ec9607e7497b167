python -m pytest tests/test_sim_gpu.py tests/test_mesh_hull_gpu.py tests/test_framework_gpu.py -x -q -m gpu 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sim_solve|sim_contacts" --csv --log-file gpurun_out/r3g_dur.csv python tools/prof_contacts.py 400 > gpurun_out/r3g_log.log 2>&1
tail -1 gpurun_out/r3g_log.log
for i in 1 2; do python bench.py --steps 200 --warmup 20 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['warm_l2']['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['gpu_launches'])"; done
