python -m pytest tests/test_framework_gpu.py -x -q -m gpu 2>&1 | tail -1
for ov in 1 0; do for f in "" "--no-flush"; do
RLOA_OVERLAP_STORE=$ov python bench.py --steps 200 --warmup 20 --no-cpu --no-extras $f | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('overlap=$ov flush=[$f]', d['ms_per_step'], d['phases_ms'], (d.get('warm_l2') or {}).get('ms_per_step'), d['gpu_launches'])"
done; done
