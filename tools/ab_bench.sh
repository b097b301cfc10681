#!/bin/bash
# Developer helper (GPU box): bench every library under build/variants/ and print one summary line each.
# usage: tools/ab_bench.sh [bench.py args...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in build/variants/*.so; do
  name=$(basename "$lib" .so)
  METROTRPL_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 "$@" \
    > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" gpurun_out/ab_$name.json <<'PY'
import json, sys
name, path = sys.argv[1:3]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d["roofline"]; s = d["stats"]
    print(f"{name:14s} value={d['value']:9.0f} e2e={d['e2e']['value']:9.0f} kernel_ms={r['kernel_ms']:7.2f} "
          f"steps/sim={s['mean_steps_per_sim']:6.1f} rej={s['mean_rejected']:.2f} frac={r['frac']:.4f} "
          f"clk={d['clocks']['sm_mhz']}")
except Exception as e:
    print(name, "FAILED", e)
PY
done
