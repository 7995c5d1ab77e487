#!/bin/bash
# round 2 evidence visit (final build): GPU suite, default bench (both arms) timed, m4 / s1pu workloads, walk probe, ncu launch
# list, full captures of K2 (100 and 1000 universes), the walk kernels, K3 and K1 — the .ncu-rep files are turned into text here
# and removed (gpurun_out is capped at 64 MiB)
tag=${1:-r2m}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
SECONDS=0
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $? after ${SECONDS}s"
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $? after ${SECONDS}s"
B="python bench.py --no-extras --no-s1 --no-cpu-baseline"
$B --workload m4 --universes 1000 --steps 3 --e2e-steps 10 > gpurun_out/${tag}_m4.log 2> gpurun_out/${tag}_m4.err; echo "m4 exit $?"
$B --workload s1pu --universes 1000 --steps 3 --e2e-steps 6 > gpurun_out/${tag}_s1pu.log 2> gpurun_out/${tag}_s1pu.err; echo "s1pu exit $?"
python tools/walk_probe.py > gpurun_out/${tag}_walk_probe.log 2>&1; echo "walk probe exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv \
    $B --steps 4 --warmup 3 --e2e-steps 4 > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list exit $?"
digest() {   # $1 = report stem: details page, raw csv, per-line stalls; the report itself is dropped
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>&1
    python tools/ncu_lines.py gpurun_out/$1.ncu-rep 45 > gpurun_out/$1_lines.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$FULL -k "regex:k2_train_universes.*512, 1, 1>" -s 8 -c 1 -o gpurun_out/${tag}_k2_100 -f \
    $B --steps 2 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_100.log 2>&1; echo "k2 100 exit $?"; digest ${tag}_k2_100
$FULL -k "regex:k2_train_universes.*512, 1, 1>" -s 2 -c 1 -o gpurun_out/${tag}_k2_1000 -f \
    $B --universes 1000 --steps 1 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_1000.log 2>&1; echo "k2 1000 exit $?"; digest ${tag}_k2_1000
$FULL -k regex:k_walk_universes -s 2 -c 1 -o gpurun_out/${tag}_walk -f \
    $B --steps 1 --warmup 3 --no-eval --e2e-steps 4 > gpurun_out/${tag}_ncu_walk.log 2>&1; echo "walk exit $?"; digest ${tag}_walk
$FULL -k regex:k_number_universes -s 2 -c 1 -o gpurun_out/${tag}_number -f \
    $B --steps 1 --warmup 3 --no-eval --e2e-steps 4 > gpurun_out/${tag}_ncu_number.log 2>&1; echo "number exit $?"; digest ${tag}_number
$FULL -k regex:k3u_energies -s 1 -c 1 -o gpurun_out/${tag}_k3u -f \
    $B --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/${tag}_ncu_k3u.log 2>&1; echo "k3u exit $?"; digest ${tag}_k3u
$FULL -k regex:k3_rank_rows -s 1 -c 1 -o gpurun_out/${tag}_k3rows -f \
    $B --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/${tag}_ncu_k3rows.log 2>&1; echo "k3rows exit $?"; digest ${tag}_k3rows
$FULL -k regex:k3_raw -c 1 -o gpurun_out/${tag}_k3raw -f \
    python tools/k3_probe.py > gpurun_out/${tag}_ncu_k3raw.log 2>&1; echo "k3raw exit $?"; digest ${tag}_k3raw
for deg in power uniform; do for opt in adagrad sgd; do
  ncu --set full --clock-control none --kernel-name-base demangled -k regex:k1_grad -s 40 -c 1 -o gpurun_out/${tag}_k1_${opt}_${deg} -f \
      python tools/bench_k1.py --opt $opt --degree $deg --steps 64 --reps 1 > gpurun_out/${tag}_ncu_k1_${opt}_${deg}.log 2>&1; echo "k1 $opt $deg exit $?"
  ncu -i gpurun_out/${tag}_k1_${opt}_${deg}.ncu-rep --page raw --csv > gpurun_out/${tag}_k1_${opt}_${deg}_raw.csv 2>&1
  ncu -i gpurun_out/${tag}_k1_${opt}_${deg}.ncu-rep --page details > gpurun_out/${tag}_k1_${opt}_${deg}_details.txt 2>&1
  rm -f gpurun_out/${tag}_k1_${opt}_${deg}.ncu-rep
done; done
du -sh gpurun_out
