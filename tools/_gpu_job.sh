set -x
cd $GRAFT_REPO_ROOT
(time python -m pytest tests/test_gpu_pairing.py tests/test_gpu_shard.py tests/test_gpu_verify.py -x -q) > gpurun_out/r02_gpu_tests_pairing.log 2>&1
tail -5 gpurun_out/r02_gpu_tests_pairing.log
python tools/extra_bench.py pairing > gpurun_out/r02_pairing_latency.jsonl 2> gpurun_out/r02_pairing_latency.err
cat gpurun_out/r02_pairing_latency.jsonl | cut -c1-400
python bench.py --power 18 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_p18_e.json 2> gpurun_out/r02_bench_p18_e.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_p18_e.json').read().strip().splitlines()[-1])
print(round(d['value']), d['legs']['contribute']['ms_per_step'], d['legs']['verify']['ms_per_step'], d['parity_spot_check'], d['verdict_all_steps'], d['roofline']['kernels_ms_verify'].get('k_same_ratio<bls12_377>'))
P
