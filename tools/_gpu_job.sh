set -x
cd $GRAFT_REPO_ROOT
(time python -m pytest tests/test_gpu_parity.py -x -q) > gpurun_out/r02_gpu_tests_parity.log 2>&1
tail -5 gpurun_out/r02_gpu_tests_parity.log
(time python bench.py) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['legs'], round(d['e2e']['value']), d['parity_spot_check'], d['verdict_all_steps'])
print(json.dumps(d['extras'])[:1500])
print(json.dumps(d['cpu_baseline'])[:600])
P
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
cut -c1-300 gpurun_out/r02_bench_reference.json
