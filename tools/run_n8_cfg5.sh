#!/bin/bash
# BASELINE configs[4] (strong scaling, 64 x 8K frames) — usage: run_n8_cfg5.sh "<torchrun Ns>" "<single-process Ns>" [cpu]
out=gpurun_out/r2_cfg5_chain_$(echo "$1$2" | tr -d ' ').jsonl
: > $out
for n in $1; do
  if [ "$n" = 1 ]; then python tools/cfg5_chain.py --steps 10 2>/dev/null | grep '^{' >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
        tools/cfg5_chain.py --steps 10 2>/dev/null | grep '^{' >> $out; fi
done
for n in $2; do
  python tools/cfg5_chain.py --single-process --gpus $n --steps 10 2>/dev/null | grep '^{' >> $out
done
[ -n "$3" ] && python tools/cfg5_chain.py --cpu 2>/dev/null | grep '^{' >> $out
cat $out | cut -c1-330
