set -x
cd $GRAFT_REPO_ROOT
(time timeout 1200 python -m pytest tests/test_gpu_mnt.py -q) > gpurun_out/r02_gpu_tests_mnt.log 2>&1
tail -30 gpurun_out/r02_gpu_tests_mnt.log
