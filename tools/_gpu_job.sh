set -x
cd $GRAFT_REPO_ROOT
(time python -m pytest tests -m gpu -q -k "bw6 or subgroup or verif or properties") > gpurun_out/r02_gpu_tests_bw6.log 2>&1
tail -4 gpurun_out/r02_gpu_tests_bw6.log
python tools/extra_bench.py 16 20 > gpurun_out/r02_extra_bw6.jsonl 2> gpurun_out/r02_extra_bw6.err
python - <<'P'
import json
for l in open('gpurun_out/r02_extra_bw6.jsonl'):
    d=json.loads(l)
    if 'verify_kernels_ms_serialised' in d:
        print(d['power'], 'contribute', round(d['contribute_powers_per_s']), 'verify', round(d['verify_powers_per_s']), d['ratio_check'], d['parity_spot_check'])
        print({k:v for k,v in d['verify_kernels_ms_serialised'].items() if 'subgroup' in k})
P
