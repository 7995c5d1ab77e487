#!/bin/bash
# quick 1-GPU verification: the GPU parity suite and one bench line
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py --no-s1 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
print('value %.1f M/s  %.2f ms/step  e2e %.1f M/s %.2f ms' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']), {k:round(v*1e3,2) for k,v in d['e2e']['host_breakdown_s_per_step'].items()}, d['e2e']['host'])"
