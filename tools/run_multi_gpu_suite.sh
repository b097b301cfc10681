#!/bin/bash
# Developer helper (8-GPU box): scaling bench + the sharded example drivers, outputs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  else $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err; fi
  python - gpurun_out/scale_n$n.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"bench n={d['n_gpus']} value={d['value']:.0f} e2e={d['e2e']['value']:.0f} ms={d['ms_per_step']:.2f}")
except Exception as e:
    print("bench FAILED", sys.argv[1], e)
PY
done
$TR --nproc-per-node 8 --master-port 29611 tools/run_dense_example.py --points 1000000 2> gpurun_out/dense_n8.err | tail -1 | tee gpurun_out/dense_n8.log
$TR --nproc-per-node 2 --master-port 29612 tools/run_dense_example.py --points 65536 2> gpurun_out/dense_n2.err | tail -1 | tee gpurun_out/dense_n2.log
$TR --nproc-per-node 8 --master-port 29613 tools/run_pt_example.py --iters 201 2> gpurun_out/pt_n8.err | tail -1 | tee gpurun_out/pt_n8.log
$TR --nproc-per-node 2 --master-port 29614 tools/run_pt_example.py --iters 201 2> gpurun_out/pt_n2.err | tail -1 | tee gpurun_out/pt_n2.log
