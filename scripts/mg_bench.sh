# usage: mg_bench.sh N [c5]  — the multi-GPU modes of bench.py on N GPUs of one box:
#   objects (C3 per rank, weak), pairs (C4, 28 view pairs sharded), shard (one mesh vertex-sharded,
#   NCCL all-reduce of the node system; with "c5" at the C5 size: 5M vertices, 20K nodes)
N=$1
run() {  # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/mg_${name}_${N}.log 2> gpurun_out/mg_${name}_${N}.err
  echo "== $name N=$N rc=$?"; tail -n 1 gpurun_out/mg_${name}_${N}.log | cut -c1-400
}
run objects --mode objects
run pairs --mode pairs
run shard --mode shard
if [ "$2" = "c5" ]; then run shard_c5 --mode shard --side 913 --nodes 20000; fi
