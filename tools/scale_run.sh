#!/bin/bash
# usage: bash tools/scale_run.sh N  -- bench at N GPUs: plain row sharding, then the row-shard x query-group grids
N=$1
port=29600
for Q in ${QS:-1 2 4}; do
  if [ $((N % Q)) -ne 0 ] || [ $Q -gt $N ]; then continue; fi
  port=$((port+1))
  out=gpurun_out/scale_g${N}_q${Q}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 30 --warmup 5 --query-groups $Q > $out.json 2> $out.err
  python - $out.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["parallelism"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print("  ", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("failed", sys.argv[1], e)
PY
done
