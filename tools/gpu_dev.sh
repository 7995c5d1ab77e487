#!/bin/bash
# development visit (K2_DEV build: d = 20 only): K2 parity tests, one-universe probes, a short bench
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "putranse_end_to_end or bernoulli_filtered or (batched_universes_match and not TransE-) or device_sampler_inside or register_resident" > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
for ids in 80 86 6; do K2_PROBE_IDS=$ids python tools/k2_probe.py 1 60 3 2>&1 | tail -1 | cut -c1-140; done
python bench.py --steps 5 --warmup 3 --no-s1 --no-cpu-baseline --no-eval > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
print('value %.1f M/s  %.2f ms/step  e2e %.1f M/s %.2f ms' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']), d['e2e']['host_breakdown_s_per_step'])"
