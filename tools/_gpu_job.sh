set -x
cd $GRAFT_REPO_ROOT
(time python -m pytest tests/test_gpu_pairing.py tests/test_gpu_shard.py tests/test_gpu_verify.py -x -q) > gpurun_out/r02_gpu_tests_pairing2.log 2>&1
tail -5 gpurun_out/r02_gpu_tests_pairing2.log
python tools/extra_bench.py pairing > gpurun_out/r02_pairing_latency2.jsonl 2> gpurun_out/r02_pairing_latency2.err
cut -c1-330 gpurun_out/r02_pairing_latency2.jsonl
