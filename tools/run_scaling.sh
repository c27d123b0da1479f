#!/bin/bash
# Multi-GPU evidence on ONE 8-GPU box (run under `gpurun --gpus 8`): strong scaling of the 64-frame sequence at 8 / 4 / 2 / 1
# ranks (BASELINE configs[3] as stated), the weak-scaling line at 8, and the data-parallel training step at 8 ranks with the
# bucketed-overlapped and the single flat gradient exchange.  One JSON line per run under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
LEAN="--no-cpu-baseline --no-gpu-reference --no-configs"
run() {  # n port outfile args...
  local n=$1 port=$2 out=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@" > $OUT/$out.json 2> $OUT/$out.err
  echo "$out rc=$? $(cut -c1-200 $OUT/$out.json | tail -1)"
}
timeout 300 python bench.py --gpus 1 --scaling strong --steps 10 --warmup 3 $LEAN > $OUT/r02_strong_n1.json 2> $OUT/r02_strong_n1.err
echo "r02_strong_n1 rc=$? $(cut -c1-200 $OUT/r02_strong_n1.json | tail -1)"
for n in 2 4 8; do
  run $n $((29500 + n)) r02_strong_n$n bench.py --gpus $n --scaling strong --steps 10 --warmup 3 $LEAN
done
# (the weak-scaling line at 1/2/4/8 is the driver's own SCALE run)
run 8 29621 r02_train_n8 tools/bench_train.py --gpus 8 --steps 30 --warmup 10
run 8 29631 r02_train_n8_flat tools/bench_train.py --gpus 8 --steps 30 --warmup 10 --bucket-mb 0
