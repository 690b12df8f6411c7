#!/bin/bash
# BASELINE.json config 2 at N GPUs: the default workload (range-sharded 2^lg-point G1 MSM + NCCL gather) for lg = 16..26, MSM only.
#   bash tools/msm_sweep_multi.sh N  > gpurun_out/msm_sweep_nN.jsonl
N=$1
for lg in 16 18 20 22 24 26; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + lg)) \
    bench.py --gpus $N --log-n $lg --no-cpu-baseline --no-l2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({k:d[k] for k in ('metric','value','unit','n_gpus','e2e','points_per_s','config','clocks','result') if k in d}))
"
done
