#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
out=gpurun_out/${tag}_k2dbg2.log
: > $out
for ids in 80 86 6; do
for cfg in "0 3" "32 3" "33 3" "0 4" "0 5" "0 6" "32 4" "32 6"; do
  set -- $cfg
  echo "ids=$ids dbg=$1 producers=$2" >> $out
  K2_PROBE_IDS=$ids PK_K2_DBG=$1 PK_K2_PRODUCERS=$2 python tools/k2_probe.py 1 60 3 2>&1 | tail -1 | cut -c1-140 >> $out
done; done
cat $out
