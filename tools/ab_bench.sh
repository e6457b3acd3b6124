#!/bin/bash
# A/B: run the contribute bench (no verify/cpu legs) for the default library and every variant
cd "$(dirname "$0")/.."
echo "== default"; python bench.py --no-cpu-baseline --no-verify --steps 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels_ms_per_step'], d['parity_spot_check'])"
for v in snark-setup_b200/csrc/variants/*.so; do
  echo "== $v"; SS_LIB=$PWD/$v python bench.py --no-cpu-baseline --no-verify --steps 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels_ms_per_step'], d['parity_spot_check'])"
done
