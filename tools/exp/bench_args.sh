i=0
for a in "--steps 7 --warmup 3 --no-cpu --no-extras --config 3" "--steps 5 --warmup 3 --no-cpu --no-extras" "--steps 3 --warmup 3 --no-cpu --no-extras --no-graph"; do
i=$((i+1))
python bench.py $a > gpurun_out/ba_$i.json 2> gpurun_out/ba_$i.err; echo "rc=$? args=$a"; tail -c 300 gpurun_out/ba_$i.json | head -c 300; echo; tail -3 gpurun_out/ba_$i.err
done
