# usage: mg_bench.sh N  — the three multi-GPU modes of bench.py on N GPUs of one box
N=$1
for mode in objects shard pairs; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --mode $mode --no-cpu-baseline --no-secondary > gpurun_out/mg_${mode}_${N}.log 2> gpurun_out/mg_${mode}_${N}.err
  echo "== $mode N=$N rc=$?"; tail -n 1 gpurun_out/mg_${mode}_${N}.log | cut -c1-700
done
