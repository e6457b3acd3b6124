set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --power 20 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_sw.json 2>/dev/null
python - <<P
import json
for l in open('gpurun_out/ab_sw.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('sw', d['value'], d['legs']['contribute']['ms_per_step'], d['legs']['verify']['ms_per_step']); print(r['kernels_ms_contribute']); print(r['kernels_ms_verify'])
P
python tools/extra_bench.py pairing
