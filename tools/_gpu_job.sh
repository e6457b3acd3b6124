set -x
cd $GRAFT_REPO_ROOT
run() { # name env...
  name=$1; shift
  for p in 18 19 20; do
    env "$@" python bench.py --power $p --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ab_msm_${name}_p$p.json 2> gpurun_out/r02_ab_msm_${name}_p$p.err
  done
}
run base SS_MSM_C_OFFSET=5
run off4 SS_MSM_C_OFFSET=4 SS_MSM_C_MAX=16
run off6 SS_MSM_C_OFFSET=6
env SS_MSM_C_OFFSET=4 SS_MSM_C_MAX=16 python -m pytest tests/test_gpu_verify.py tests/test_gpu_shard.py -q 2>&1 | tail -2
python - <<'P'
import json
for name in ('base','off4','off6'):
  for p in (18,19,20):
    try:
      d=json.loads(open(f'gpurun_out/r02_ab_msm_{name}_p{p}.json').read().strip().splitlines()[-1])
      kv=d['roofline']['kernels_ms_verify']
      print(name, p, round(d['legs']['verify']['ms_per_step'],2), 'acc', kv.get('k_msm_accumulate<bls12_377.g1>'), kv.get('k_msm_accumulate<bls12_377.g2>'), 'red', kv.get('k_msm_reduce<bls12_377.g1>'), kv.get('k_msm_reduce<bls12_377.g2>'), 'sort', kv.get('k_msm_sort<bls12_377.g1>'), d['verdict_all_steps'])
    except Exception as e:
      print(name, p, 'ERR', e)
P
