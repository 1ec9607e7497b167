python -m pytest tests/test_naf_learn_cluster_gpu.py tests/test_grad_exchange_gpu.py tests/test_framework_gpu.py tests/test_zz_learning_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/learn_cluster_profile.py 1024 2>&1 | egrep "adam|total|final" 
python tools/learn_cluster_profile.py 1024 flush 2>&1 | egrep "adam|total|final"
python tools/learn_timing.py 1024
python bench.py --steps 200 --warmup 20 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['warm_l2']['ms_per_step'], d['phases_ms']['sample_learn'])"
