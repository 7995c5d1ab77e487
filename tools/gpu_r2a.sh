#!/bin/bash
# round 2, first visit: GPU parity suite (incl. the new full-length and closed-form tests) + launch pipelining probe
tag=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python tools/pipeline_probe.py > gpurun_out/${tag}_pipeline.log 2> gpurun_out/${tag}_pipeline.err
grep "^n=" gpurun_out/${tag}_pipeline.log
