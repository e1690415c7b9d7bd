#!/bin/bash
# usage: run_nN.sh N   (2 or 4): the N-rank bench line
N=$1
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_n$N.json 2> $O/bench_n$N.err
tail -2 $O/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read())
print("N=$N: value",d["value"],"ms_per_step",d["ms_per_step"])
print("e2e",d["e2e"]["value"], "depth", d["e2e"]["depth_input"]["value"], "depth idx", d["e2e"]["depth_input_idx_only"]["value"])
print("closed", {k:round(v.get("frames_per_s")) for k,v in d["e2e_closed_loop"].items()})
print("batched", d["batched_sequences"]["frames_per_s"], "nn", d["nn"]["query_ms"], d["nn"]["build_ms"], d["nn"].get("weak_scaling"))
PY
