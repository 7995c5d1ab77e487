#!/bin/bash
tag=${1:-run}
K2_PROBE_IDS=86 PK_K2_DBG=32 ncu --set full --clock-control none --import-source on -k regex:k2_train_universes -s 1 -c 1 -o gpurun_out/${tag}_prod -f \
    python tools/k2_probe.py 1 60 1 > gpurun_out/${tag}_ncu_prod.log 2>&1
tail -2 gpurun_out/${tag}_ncu_prod.log
