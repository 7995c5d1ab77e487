#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, one full capture of the K2 kernel,
# per-universe times.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $?"
python tools/k2_universe_times.py 100 > gpurun_out/${tag}_k2_universe_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-s1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k2_train_universes.*512, .int.1, .int.1>' -s 2 -c 1 -o gpurun_out/${tag}_k2 -f \
    python bench.py --steps 1 --warmup 3 --no-s1 --no-cpu-baseline --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2.log 2>&1
ls -la gpurun_out | tail -12
