#!/bin/bash
# dev-build visit: K2 parity tests that use d = 20, smoke, the timing experiments, a short bench
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "putranse_end_to_end or bernoulli_filtered or (batched_universes_match and not TransE) or device_sampler_inside" > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_k2dbg.sh $tag
python bench.py --steps 5 --warmup 3 --no-s1 --no-cpu-baseline --no-eval > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
print('value %.1f M/s  %.2f ms/step  e2e %.1f M/s %.2f ms' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"
