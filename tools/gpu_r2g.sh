#!/bin/bash
# quick: GPU suite with failure details, fb15k parity probe
tag=${1:-r2g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
grep -n "^E  " gpurun_out/${tag}_pytest.log | head -30
python tools/parity_probe.py fb15k > gpurun_out/${tag}_parity_fb15k.log 2>gpurun_out/${tag}_parity_fb15k.err; tail -22 gpurun_out/${tag}_parity_fb15k.log; tail -5 gpurun_out/${tag}_parity_fb15k.err
python tools/pipeline_probe.py 2>/dev/null | grep "^n= 100"
