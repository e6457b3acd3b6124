set -x
cd $GRAFT_REPO_ROOT
for p in 18 20; do for pr in 0 1; do
SS_PRIORITY_LANES=$pr python bench.py --power $p --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ab_prio${pr}_p$p.json 2> gpurun_out/r02_ab_prio${pr}_p$p.err
done; done
python - <<'P'
import json
for p in (18,20):
  for pr in (0,1):
    d=json.loads(open(f'gpurun_out/r02_ab_prio{pr}_p{p}.json').read().strip().splitlines()[-1])
    print(p, pr, round(d['value']), round(d['legs']['contribute']['ms_per_step'],2), round(d['legs']['verify']['ms_per_step'],2), d['parity_spot_check'], d['verdict_all_steps'])
P
python -m pytest tests/test_gpu_verify.py tests/test_gpu_shard.py -q 2>&1 | tail -2
