#!/bin/bash
tag=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
python tools/parity_probe.py > gpurun_out/${tag}_parity.log 2> gpurun_out/${tag}_parity.err; tail -45 gpurun_out/${tag}_parity.log
python bench.py --workload m4 --universes 1000 --steps 3 --e2e-steps 3 --no-extras --no-s1 --no-cpu-baseline > gpurun_out/${tag}_m4.log 2> gpurun_out/${tag}_m4.err; echo "m4 exit $?"; cut -c1-1200 gpurun_out/${tag}_m4.log; tail -5 gpurun_out/${tag}_m4.err
python bench.py --model transd --steps 8 --no-extras --no-s1 --no-cpu-baseline --no-eval > gpurun_out/${tag}_transd.log 2> gpurun_out/${tag}_transd.err; echo "transd exit $?"; cut -c1-400 gpurun_out/${tag}_transd.log
