#!/bin/bash
# ncu full capture of the STAGED K2 launch inside bench.py (launch order per call: unstaged class, staged class)
tag=${1:-run}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k2_train_universes.*512, .int.1, .int.1>' -s 2 -c 1 -o gpurun_out/${tag}_k2 -f \
    python bench.py --steps 1 --warmup 3 --no-s1 --no-cpu-baseline --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2.log 2>&1
K2_PROBE_IDS=80 python tools/k2_probe.py 1 60 3 > gpurun_out/${tag}_probe80.log 2>&1
K2_PROBE_IDS=80 ncu --set full --clock-control none --import-source on -k regex:k2_train_universes -s 1 -c 1 -o gpurun_out/${tag}_k2u80 -f \
    python tools/k2_probe.py 1 60 1 > gpurun_out/${tag}_ncu_k2u80.log 2>&1
cat gpurun_out/${tag}_probe80.log
