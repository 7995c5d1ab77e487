#!/bin/bash
# 2-GPU visit: two-rank evaluation/checkpoint test, 2-GPU bench (m2 weak, m4 strong), eval probe on one GPU
tag=${1:-r2d}
mkdir -p gpurun_out
python -m pytest tests/test_product.py tests/test_parity_full.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench2.log 2> gpurun_out/${tag}_bench2.err; echo "bench2 exit $?"; cut -c1-300 gpurun_out/${tag}_bench2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload m4 --universes 1000 --steps 3 --e2e-steps 3 > gpurun_out/${tag}_m4_2.log 2> gpurun_out/${tag}_m4_2.err; echo "m4x2 exit $?"; cut -c1-300 gpurun_out/${tag}_m4_2.log
python tools/eval_probe.py 400 > gpurun_out/${tag}_eval_probe.log 2>&1; head -8 gpurun_out/${tag}_eval_probe.log; grep -A25 "cumulative" gpurun_out/${tag}_eval_probe.log | cut -c1-150 | head -30
