#!/bin/bash
# multi-GPU bench line under torchrun: bash scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-r2m}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_${TAG}_n$N.json") if l.startswith("{")][-1]
print("value", d["value"], "ms", d["ms_per_step"])
for k in ("strong","train_dp"):
    print(k, json.dumps(d.get(k))[:1800])
print("vae", d.get("vae",{}).get("value"), "act", d.get("act",{}).get("act_ms"))
PY
