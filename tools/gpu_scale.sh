#!/bin/bash
# weak-scaling bench on one box: N = 8, 4, 2 ranks (torchrun, NCCL), one JSON line each
mkdir -p gpurun_out
for n in ${NS:-8 4 2}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r1_bench_${n}gpu.log 2> gpurun_out/r1_bench_${n}gpu.err
  echo "N=$n exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r1_bench_${n}gpu.log') if l.startswith('{')][-1])
    print('N=%d value %.1f M/s %.2f ms/step e2e %.1f M/s eval %s' % (d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d.get('eval',{}).get('filtered')))
except Exception as e: print('parse failed', e)
PY
done
