#!/bin/bash
# A/B: run the bench for the default library and every variant under snark-setup_b200/csrc/variants/
cd "$(dirname "$0")/.."
summ='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("contribute", round(d["value"]), round(d["ms_per_step"],1), "serial", round(d["roofline"]["serialised_step_ms"],1), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernels_ms_per_step"], d["parity_spot_check"], "frac", round(d["roofline"]["frac"],3)); v=d.get("verify"); print("verify", round(v["value"]), round(v["ms_per_step"],1), v["ratio_and_reemit_check"], v["kernels_ms_per_step_serialised"]) if v else None'
echo "== default"; python bench.py --no-cpu-baseline --steps 3 2>&1 | python -c "$summ"
for v in snark-setup_b200/csrc/variants/*.so; do
  [ -e "$v" ] || continue
  echo "== $v"; SS_LIB=$PWD/$v python bench.py --no-cpu-baseline --steps 3 2>&1 | python -c "$summ"
done
