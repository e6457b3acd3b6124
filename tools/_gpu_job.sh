set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gpu_tests_a.log 2>&1
tail -5 gpurun_out/r02_gpu_tests_a.log
python bench.py --power 18 --steps 3 --warmup 3 > gpurun_out/r02_bench_p18.json 2> gpurun_out/r02_bench_p18.err
tail -c 3000 gpurun_out/r02_bench_p18.json; tail -5 gpurun_out/r02_bench_p18.err
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r02_bench_p22_a.json 2> gpurun_out/r02_bench_p22_a.err
tail -c 6000 gpurun_out/r02_bench_p22_a.json; tail -5 gpurun_out/r02_bench_p22_a.err
