#!/bin/bash
# r02 final-code check on one 8-GPU box: N = 8 with 2 and 3 batches in flight, N = 1
mkdir -p gpurun_out
run() {  # name, gpus, extra args
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
        bench.py --gpus $n --steps 30 --warmup 5 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  python - gpurun_out/$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["parallelism"], "in-flight", d["config"]["batches_in_flight"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3),
          "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["result_digest"][:12],
          d["id_parity"] and d["id_parity"]["bit_exact_rate"], d["config"]["dense_path"]["n_segments"])
    print("  ", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("failed", sys.argv[1], e)
PY
}
run u8 8
run u8_f3 8 --in-flight 3
run u8_f1 8 --in-flight 1
run u1 1 --no-c1 --sparse-c4-docs 0 --no-side-configs
