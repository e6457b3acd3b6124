set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_pairing.py -q -x > gpurun_out/r02_gpu_tests_pairing_units.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_pairing_units.log
python tools/extra_bench.py pairing > gpurun_out/pairing_latency_units.jsonl; cat gpurun_out/pairing_latency_units.jsonl
