#!/bin/bash
# the driver's own commands at N ranks (both arms, default flags) + smoke()
n=${1:-2}; tag=${2:-r2u}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 \
    > gpurun_out/${tag}_bench_${n}gpu.log 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "ours exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 2 --warmup 1 \
    > gpurun_out/${tag}_ref_${n}gpu.log 2> gpurun_out/${tag}_ref_${n}gpu.err; echo "reference exit $?"
N=$n TAG=$tag python - <<'P'
import json, os
for w in ("bench", "ref"):
    f = "gpurun_out/%s_%s_%sgpu.log" % (os.environ["TAG"], w, os.environ["N"])
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(w, d.get("value"), d.get("ms_per_step"), json.dumps(d.get("e2e"))[:400])
            print("   eval", json.dumps(d.get("eval"))[:700])
P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
