#!/bin/bash
# K3 with register rows against the local-memory instantiation; eval probe; GPU suite
tag=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
echo "--- register rows"; python tools/eval_probe.py 400 2>&1 | grep "^universes"; python tools/k3_probe.py 2>&1 | tail -4
echo "--- PK_K3_ROWS=256 (local memory rows)"; PK_K3_ROWS=256 python tools/eval_probe.py 400 2>&1 | grep "^universes"; PK_K3_ROWS=256 python tools/k3_probe.py 2>&1 | tail -4
