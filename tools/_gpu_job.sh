set -x
cd $GRAFT_REPO_ROOT
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gpu_tests_c.log 2>&1
tail -5 gpurun_out/r02_gpu_tests_c.log
python bench.py --power 20 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ab_pairfull_p20.json 2> gpurun_out/r02_ab_pairfull_p20.err
SS_LIB=$PWD/snark-setup_b200/csrc/variants/libss_g2l4.so python bench.py --power 20 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ab_pairfull_minb4_p20.json 2> gpurun_out/r02_ab_pairfull_minb4_p20.err
SS_LIB=$PWD/snark-setup_b200/csrc/variants/libss_g2l2.so python bench.py --power 20 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ab_pairfull_minb2_p20.json 2> gpurun_out/r02_ab_pairfull_minb2_p20.err
python - <<'P'
import json
for f in ('r02_ab_pairfull_p20','r02_ab_pairfull_minb4_p20','r02_ab_pairfull_minb2_p20'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), d['legs']['contribute']['ms_per_step'], d['legs']['verify']['ms_per_step'], d['parity_spot_check'], d['verdict_all_steps'])
        print(' C', {k:v for k,v in d['roofline']['kernels_ms_contribute'].items() if 'g2' in k})
        print(' V', {k:v for k,v in d['roofline']['kernels_ms_verify'].items() if 'g2' in k})
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/'+f+'.err').read()[-1500:])
P
