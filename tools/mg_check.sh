# N-GPU check of bench.py: NCCL all-reduce vs peer-memory exchange, repeated.  Usage: bash tools/mg_check.sh <ngpus> <reps>
n=${1:-2}; reps=${2:-2}
mkdir -p gpurun_out
for i in $(seq 1 $reps); do
  for mode in nccl peer; do
    RLOA_GRAD_EXCHANGE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/mg_${mode}_n${n}.out 2> gpurun_out/mg_${mode}_n${n}.err
    echo "$mode n=$n rc=$? $(python -c "
import json
d=json.loads(open('gpurun_out/mg_${mode}_n${n}.out').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],5), d['phases_ms']['sample_learn'], d['graphed'], d['graph_error'])")"
  done
done
