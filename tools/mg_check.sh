mkdir -p gpurun_out
for mode in ""; do
  tag=$(echo "x$mode" | tr -d ' -')
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM%9)) bench.py --gpus 2 --steps 60 --warmup 6 $mode > gpurun_out/mg_$tag.out 2> gpurun_out/mg_$tag.err
  echo "mode=[$mode] rc=$?"; tail -c 1500 gpurun_out/mg_$tag.out; grep -v "^$" gpurun_out/mg_$tag.err | tail -8
  sleep 3
done
