#!/bin/bash
# round 2 visit: GPU parity suite, bench (both arms), ncu launch list
tag=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $?"
cut -c1-3000 gpurun_out/${tag}_bench.log
