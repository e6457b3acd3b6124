set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x -k "multi_tile or parity or properties or shard" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/pageable_after.json 2> gpurun_out/pageable_after.err
python - <<P
import json
for l in open('gpurun_out/pageable_after.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['e2e']['pageable'])
P
