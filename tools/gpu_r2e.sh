#!/bin/bash
# round 2 evidence visit: full bench (both arms), s1pu workload, ncu launch list, full captures of K2 (100 and 1000 universes) and K3
tag=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $?"
python bench.py --workload s1pu --universes 1000 --steps 3 --e2e-steps 3 --no-extras --no-s1 --no-cpu-baseline > gpurun_out/${tag}_s1pu.log 2> gpurun_out/${tag}_s1pu.err; echo "s1pu exit $?"; cut -c1-200 gpurun_out/${tag}_s1pu.log
B="python bench.py --no-extras --no-s1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    $B --steps 4 --warmup 3 --e2e-steps 4 > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k2_train_universes -s 8 -c 1 -o gpurun_out/${tag}_k2_100 -f \
    $B --steps 2 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_100.log 2>&1; echo "k2 100 exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k2_train_universes -s 4 -c 1 -o gpurun_out/${tag}_k2_1000 -f \
    $B --universes 1000 --steps 1 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_1000.log 2>&1; echo "k2 1000 exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k3u_energies -s 1 -c 1 -o gpurun_out/${tag}_k3u -f \
    $B --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/${tag}_ncu_k3u.log 2>&1; echo "k3u exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k3_rank_rows -s 1 -c 1 -o gpurun_out/${tag}_k3rows -f \
    $B --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/${tag}_ncu_k3rows.log 2>&1; echo "k3rows exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k3_raw -c 1 -o gpurun_out/${tag}_k3raw -f \
    python tools/k3_probe.py > gpurun_out/${tag}_ncu_k3raw.log 2>&1; echo "k3raw exit $?"
ls -la gpurun_out | grep ${tag}
