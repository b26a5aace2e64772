#!/bin/bash
# order in which the encoder-phase gradient buckets are all-reduced (2 GPUs): ms/step per order
p=29560
for o in "embed,enc_a,enc_v,enc_t_l2,enc_t" "embed,enc_a,enc_t_l2,enc_v,enc_t" "enc_a,embed,enc_v,enc_t_l2,enc_t" "embed,enc_t_l2,enc_a,enc_v,enc_t" "enc_a,enc_v,embed,enc_t_l2,enc_t"; do
  p=$((p+1))
  r=$(MMDA_AR_ORDER=$o timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
  echo "$o -> $r"
done
