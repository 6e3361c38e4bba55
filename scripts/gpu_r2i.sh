#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_train_check.py 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/dp_train_check_r2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2i_n2.json 2> gpurun_out/bench_r2i_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench_r2i_n2.json"):
    if l.startswith("{"): d=json.loads(l)
print("value", d["value"], "ms", d["ms_per_step"])
print(json.dumps(d.get("train_dp"))[:1500])
print("vae", d.get("vae",{}).get("value"), "act", d.get("act",{}).get("act_ms"), d.get("act",{}).get("idm_loop_ms"))
PY
