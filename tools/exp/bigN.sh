python tools/bench_sim_quick.py panda 131072 2>&1 | grep "step "
RLOA_CONTACTS=0 python tools/bench_sim_quick.py panda 131072 2>&1 | grep "step "
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sim_" -s 220 -c 8 --csv --log-file gpurun_out/r3j_big.csv python tools/bench_sim_quick.py panda 131072 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r3j_big.csv')) if len(r)>5]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    print(r[ix['Kernel Name']][:40], r[ix['Metric Name']], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
