#!/bin/bash
# timing experiments: PK_K2_DBG bit switches (see train_universes.cu), universe 80 (B=99) and 86 (B=45)
tag=${1:-run}
mkdir -p gpurun_out
out=gpurun_out/${tag}_k2dbg.log
: > $out
for ids in 80 86 6; do
for dbg in ${DBGS:-0 1 2 3 4 8 16 31}; do
  echo "ids=$ids dbg=$dbg" >> $out
  K2_PROBE_IDS=$ids PK_K2_DBG=$dbg python tools/k2_probe.py 1 60 3 2>&1 | tail -1 | cut -c1-140 >> $out
done; done
cat $out
