#!/bin/bash
tag=${1:-w1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_walk_device.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/${tag}_pytest.log
timeout 300 python tools/walk_probe.py > gpurun_out/${tag}_probe.log 2>&1; echo "probe exit $?"; tail -12 gpurun_out/${tag}_probe.log
