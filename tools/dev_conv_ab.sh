#!/bin/bash
# Developer helper (GPU box): self-convergence statistics (tools/dev_conv.py) for every variant library
cd "$(dirname "$0")/.."
for lib in build/variants/*.so; do echo "== $lib"; METROTRPL_B200_LIB=$lib timeout 300 python tools/dev_conv.py 2>&1 | grep -E "7v9|set |percentile" | head -5; done
