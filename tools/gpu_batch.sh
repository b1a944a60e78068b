#!/bin/bash
# scratch: emulate the host-core budget of 8 ranks on 16 cores with 2 ranks on 4 cores, sweep the narrowing workers
for t in 1 2 3; do
  echo "== 4 cores, IAM_HOST_THREADS=$t"
  IAM_HOST_THREADS=$t taskset -c 0-3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$t \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-spot 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print(round(d['value']), round(e['value']), round(e['ms_per_step'],1), e['frames_narrowed_on_host'], e['h2d_bytes_per_step'], e['timeline_ms_rank0'])"
done
