#!/bin/bash
# round 2 scaling visit on one multi-GPU box: (optionally) the two-rank product test, then bench.py under torchrun for each N
# usage: gpu_scale2.sh TAG "8 4 2" [m2|m4|s1pu] [extra bench args]
tag=${1:-r2s}; NS=${2:-"8"}; wl=${3:-m2}; shift 3
mkdir -p gpurun_out
if [ -n "$TWO_RANK_TEST" ]; then
  timeout 600 python -m pytest tests/test_product.py -m gpu -q -k two_rank > gpurun_out/${tag}_two_rank.log 2>&1; echo "two-rank test exit $?"; tail -3 gpurun_out/${tag}_two_rank.log
fi
for n in $NS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --workload $wl --steps 10 --warmup 3 --e2e-steps 30 --no-extras --no-s1 --no-cpu-baseline "$@" \
      > gpurun_out/${tag}_${wl}_${n}gpu.log 2> gpurun_out/${tag}_${wl}_${n}gpu.err
  echo "N=$n exit $?"
  N=$n TAG=$tag WL=$wl python - <<'PY'
import json, os
f = 'gpurun_out/%s_%s_%sgpu.log' % (os.environ['TAG'], os.environ['WL'], os.environ['N'])
try:
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print('N=%d value %.1f M/s %.2f ms/step | e2e %.1f M/s %.2f ms/step | host %s | eval %s' % (
        d['n_gpus'], d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6, d['e2e']['ms_per_step'],
        {k: round(v * 1e3, 2) for k, v in d['e2e'].get('host_breakdown_s_per_step', {}).items()}, d.get('eval')))
    print('   host', d['e2e'].get('host'))
except Exception as e:
    print('parse failed', e)
PY
done
