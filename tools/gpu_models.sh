#!/bin/bash
# the other configurations of BASELINE.json on one GPU: PuTransH / PuTransD (configs[2]), C1 (configs[0]), ranking probe
tag=${1:-run}
mkdir -p gpurun_out
for m in transh transd; do
  python bench.py --model $m --steps 5 --warmup 3 --no-s1 --no-cpu-baseline > gpurun_out/${tag}_bench_$m.log 2> gpurun_out/${tag}_bench_$m.err
  python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_$m.log').read().strip().splitlines()[-1])
print('$m value %.1f M/s  %.2f ms/step  e2e %.1f M/s %.2f ms  eval %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d.get('eval',{}).get('filtered')))"
done
python tools/bench_c1.py > gpurun_out/${tag}_c1.log 2>&1; tail -3 gpurun_out/${tag}_c1.log
python tools/k3_probe.py > gpurun_out/${tag}_k3.log 2>&1; tail -3 gpurun_out/${tag}_k3.log
python tools/bench_k1.py --opt sgd --steps 100 --reps 3 > gpurun_out/${tag}_k1_sgd.log 2>&1; tail -1 gpurun_out/${tag}_k1_sgd.log | cut -c1-400
