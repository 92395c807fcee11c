#!/bin/bash
# 8-GPU scaling check (run under gpurun --gpus 8): N=1 then N=8 of the default workload; "1024" as first argument
# adds the 1024-cell workload (8192 chains per GPU)
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s1.json 2>gpurun_out/s1.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s8_1.json 2>gpurun_out/s8_1.err
FILES="s1 s8_1"
if [ "$1" = "1024" ]; then
  timeout 150 python bench.py --workload burgers_pcn_1024 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s1_1024.json 2>gpurun_out/s1_1024.err
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29507 bench.py --gpus 8 --workload burgers_pcn_1024 --steps 3 --warmup 3 > gpurun_out/s8_1024.json 2>gpurun_out/s8_1024.err
  FILES="$FILES s1_1024 s8_1024"
fi
python - $FILES <<PY
import json, sys
for f in sys.argv[1:]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],3), "allreduce", round(d["allreduce_ms"],3), "frac", round(d["roofline"]["frac"],4))
    for r in d["per_rank_ms"]["rows"]: print("   ", r)
PY
