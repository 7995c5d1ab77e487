#!/bin/bash
# device-walk integration visit: GPU suite, pipeline probe (host breakdown), bench m2 + m4 + s1pu end to end
tag=${1:-r2h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
grep -n "^E  " gpurun_out/${tag}_pytest.log | head -30
python tools/pipeline_probe.py 2>gpurun_out/${tag}_pipeline.err | grep "^n=" | tee gpurun_out/${tag}_pipeline.log
python bench.py --steps 20 --warmup 5 --no-s1 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
TAG=${tag} python - <<'P'
import json, os
for line in open('gpurun_out/%s_bench.log' % os.environ['TAG']):
    if line.startswith('{'):
        d = json.loads(line)
        print('value %.3f G/s  e2e %.3f G/s  e2e ms/step %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step']))
        print(d['e2e'].get('host_breakdown_s_per_step'))
        for k, v in (d.get('extras') or {}).items():
            print(k, {a: b for a, b in v.items() if a in ('value', 'e2e_value', 'ms_per_step', 'e2e_ms_per_step', 'ms_per_call')})
        print('eval', d.get('eval'))
P
