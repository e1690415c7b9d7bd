#!/bin/bash
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_n8.json 2> $O/bench_n8.err
tail -2 $O/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 profiles/prof_pcie_multi.py > $O/r2_pcie_n8.txt 2> $O/pcie_n8.err
cat $O/r2_pcie_n8.txt
nvidia-smi topo -m > $O/r2_topo.txt 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n8.json"))
print("value",d["value"],"ms_per_step",d["ms_per_step"])
print("spread",d["timed_region"]["ms_per_step_min_median_max_per_rank"])
print("e2e",d["e2e"]["value"], "depth", d["e2e"]["depth_input"]["value"])
print("closed", {k:v.get("frames_per_s") for k,v in d["e2e_closed_loop"].items()} if isinstance(d.get("e2e_closed_loop"),dict) else d.get("e2e_closed_loop"))
print("nn", json.dumps(d["nn"])[:1500])
PY
