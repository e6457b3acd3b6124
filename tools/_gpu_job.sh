set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tee gpurun_out/r02_gpu_multi_tests_2gpu_final.log | tail -3
