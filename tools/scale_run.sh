#!/bin/bash
# Runs under `gpurun --gpus N`: the judged line on N independent banks and the one-stream form on N GPUs; JSON lines into gpurun_out/.
# Usage: tools/scale_run.sh N [modes]
N=${1:-2}
MODES=${2:-"bands channels"}
mkdir -p gpurun_out
for mode in $MODES; do
    if [ "$N" = "1" ]; then
        timeout 400 python bench.py --steps 60 --warmup 5 --shard $mode 2>gpurun_out/n${N}_$mode.err | tail -1 > gpurun_out/n${N}_$mode.json
    else
        timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N bench.py \
            --gpus $N --steps 60 --warmup 5 --shard $mode 2>gpurun_out/n${N}_$mode.err | tail -1 > gpurun_out/n${N}_$mode.json
    fi
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/n${N}_$mode.json"))
    print("N=$N $mode: device %.0f MS/s (%.3f ms/step)  e2e %.0f MS/s (%.3f ms/step)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["step_detail"])
except Exception as e:
    print("N=$N $mode failed:", e)
    print(open("gpurun_out/n${N}_$mode.err").read()[-1500:])
PY
done
