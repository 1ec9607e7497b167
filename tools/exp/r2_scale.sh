# usage: bash tools/exp/r2_scale.sh N tag     (under gpurun --gpus N)
N=$1; tag=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 200 --warmup 20 --no-extras > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
tail -c 300 gpurun_out/${tag}_bench_n$N.json; tail -3 gpurun_out/${tag}_bench_n$N.err
RLOA_GRAD_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 200 --warmup 20 --no-extras --no-cpu > gpurun_out/${tag}_bench_n${N}_nccl.json 2> gpurun_out/${tag}_bench_n${N}_nccl.err
tail -c 200 gpurun_out/${tag}_bench_n${N}_nccl.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/learn_cluster_profile.py 1024 > gpurun_out/${tag}_learn_timeline_n$N.txt 2>&1
grep -A24 "main cluster" gpurun_out/${tag}_learn_timeline_n$N.txt | head -30
