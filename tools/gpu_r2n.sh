#!/bin/bash
# validation of a new build: GPU suite, default bench timed, m4 and s1pu end to end
tag=${1:-r2n}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log; grep -n "^E  " gpurun_out/${tag}_pytest.log | head -20
SECONDS=0
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $? after ${SECONDS}s"
B="python bench.py --no-extras --no-s1 --no-cpu-baseline --no-eval"
$B --workload m4 --universes 1000 --steps 3 --e2e-steps 8 > gpurun_out/${tag}_m4.log 2> gpurun_out/${tag}_m4.err; echo "m4 exit $?"
$B --workload s1pu --universes 1000 --steps 3 --e2e-steps 8 > gpurun_out/${tag}_s1pu.log 2> gpurun_out/${tag}_s1pu.err; echo "s1pu exit $?"
TAG=${tag} python - <<'P'
import json, os
for w in ("bench", "m4", "s1pu"):
    for line in open('gpurun_out/%s_%s.log' % (os.environ['TAG'], w)):
        if line.startswith('{'):
            d = json.loads(line)
            print(w, 'value %.3f G/s %.2f ms | e2e %.3f G/s %.2f ms | host %s' % (d['value'] / 1e9, d['ms_per_step'], d['e2e']['value'] / 1e9, d['e2e']['ms_per_step'],
                  {k: round(v * 1e3, 2) for k, v in d['e2e']['host_breakdown_s_per_step'].items()}))
            for k, v in (d.get('extras') or {}).items():
                print('   ', k, {a: (round(b / 1e9, 3) if 'value' in a else round(b, 2)) for a, b in v.items() if a in ('value', 'e2e_value', 'ms_per_step', 'e2e_ms_per_step', 'ms_per_call', 'error')})
            if 'eval' in d: print('    eval', {k: v for k, v in d['eval'].items() if k != 'filtered'})
P
