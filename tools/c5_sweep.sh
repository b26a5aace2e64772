#!/bin/bash
# BASELINE configs[4]: batch-sharded data parallel, global batch 4096, sequence-length sweep.
#   tools/c5_sweep.sh <n_gpus> "<seq lengths>"     -> gpurun_out/r02_c5_n<N>_t<T>.json
N=$1; SEQS=$2; PB=$((4096 / N)); mkdir -p gpurun_out
for T in $SEQS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29600 + T % 97)) bench.py --gpus $N --steps 8 --warmup 3 --no-cpu --batch $PB --seq $T \
    2> gpurun_out/c5_n${N}_t${T}.err | tail -1 > gpurun_out/r02_c5_n${N}_t${T}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_c5_n${N}_t${T}.json"))
    print("N=${N} T=${T} B/GPU=${PB}:", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "samples/s e2e",
          round(d["e2e"]["value"]), "graph", d["cuda_graph"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          "dp_check", d["dp_check"]["grad_max_err_over_max_vs_single_gpu"], d["dp_check"]["loss_max_rel_vs_cpu_oracle"])
except Exception as e:
    print("N=${N} T=${T} FAILED", e); print(open("gpurun_out/c5_n${N}_t${T}.err").read()[-1500:])
PY
done
