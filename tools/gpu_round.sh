mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c_pytest.log
tail -15 gpurun_out/r1c_pytest.log
for mode in "--no-graph" "" "--trunk tc"; do
  timeout 300 python bench.py --no-cpu $mode > gpurun_out/r1c_bench_$(echo $mode | tr -d ' -').json 2> gpurun_out/r1c_bench.err; tail -3 gpurun_out/r1c_bench.err; cat gpurun_out/r1c_bench_$(echo $mode | tr -d ' -').json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-graph > gpurun_out/r1c_ncu_bench.log 2>&1
