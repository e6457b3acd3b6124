set -x
cd $GRAFT_REPO_ROOT
( time SS_TEST_FULL=1 python -m pytest tests/test_gpu_properties.py -m gpu -q -x ) > gpurun_out/r02_gpu_tests_full_sizes.log 2>&1; tail -6 gpurun_out/r02_gpu_tests_full_sizes.log
for m in 2 6; do
SS_PAIR_KERNELS=$m python bench.py --power 20 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_pair$m.json 2>/dev/null
python - <<P
import json
for l in open('gpurun_out/ab_pair$m.json'):
    if l.startswith('{'):
        d=json.loads(l); kv=d['roofline']['kernels_ms_verify']; print('mask $m', d['legs']['verify']['ms_per_step'], {k:v for k,v in kv.items() if 'g2' in k})
P
done
