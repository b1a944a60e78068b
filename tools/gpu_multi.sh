#!/bin/bash
# N-GPU check (gpurun --gpus N): the driver's launch line for N ranks -- BASELINE configs[3], pair-sharded, with the e2e leg.
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/host_n$N.txt; free -g | head -2 >> gpurun_out/host_n$N.txt
for n in $N ${ALSO_N}; do
  echo "== bench --gpus $n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 2>&1 | tee gpurun_out/bench_n$n.log | tail -1 | cut -c1-5000
done
${EXTRA_CMD:-true}
