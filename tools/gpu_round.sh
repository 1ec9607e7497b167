# One GPU-box round: parity tests, bench (default + tensor-core trunk), launch list.  Usage: bash tools/gpu_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --trunk fp32 > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --trunk tc > gpurun_out/${tag}_bench_tc.json 2>> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
for f in fp32 tc; do python -c "import sys,json; d=json.loads(open('gpurun_out/${tag}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['phases_ms'], d['graphed'], d['graph_error'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-graph --trunk tc > gpurun_out/${tag}_ncu_bench.log 2>&1
