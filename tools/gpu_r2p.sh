#!/bin/bash
# final-build evidence: GPU suite, default bench (both arms), ncu launch list, full captures of K2 (100- and 1000-universe launches)
tag=${1:-r2p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
SECONDS=0
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $? after ${SECONDS}s"
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $? after ${SECONDS}s"
B="python bench.py --no-extras --no-s1 --no-cpu-baseline"
$B --workload m4 --universes 1000 --steps 3 --e2e-steps 10 --no-eval > gpurun_out/${tag}_m4.log 2> gpurun_out/${tag}_m4.err; echo "m4 exit $?"
$B --workload s1pu --universes 1000 --steps 3 --e2e-steps 10 --no-eval > gpurun_out/${tag}_s1pu.log 2> gpurun_out/${tag}_s1pu.err; echo "s1pu exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
    $B --steps 4 --warmup 3 --e2e-steps 4 > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list exit $?"
digest() {
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>&1
    python tools/ncu_lines.py gpurun_out/$1.ncu-rep 45 > gpurun_out/$1_lines.txt 2>&1
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$FULL -k 'regex:k2_train_universes.*\)512, \(int\)1, \(int\)1>' -s 8 -c 1 -o gpurun_out/${tag}_k2_100 -f \
    $B --steps 2 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_100.log 2>&1; echo "k2 100 exit $?"; digest ${tag}_k2_100
$FULL -k 'regex:k2_train_universes.*\)512, \(int\)1, \(int\)1>' -s 2 -c 1 -o gpurun_out/${tag}_k2_1000 -f \
    $B --universes 1000 --steps 1 --warmup 3 --no-eval --e2e-steps 1 > gpurun_out/${tag}_ncu_k2_1000.log 2>&1; echo "k2 1000 exit $?"; digest ${tag}_k2_1000
# K1 whole steps in steady state, caches NOT flushed between kernels: DRAM bytes per step against the algorithmic figure
for opt in adagrad sgd; do
  ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 60 -c 40 --csv \
      --log-file gpurun_out/${tag}_k1_steps_${opt}.csv python tools/bench_k1.py --opt $opt --degree uniform --steps 64 --reps 2 > gpurun_out/${tag}_ncu_k1_steps_${opt}.log 2>&1; echo "k1 steps $opt exit $?"
done
du -sh gpurun_out
