set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name --format=csv
(time python -m pytest tests/test_gpu_multi.py -x -q) > gpurun_out/r02_gpu_multi_tests.log 2>&1
tail -5 gpurun_out/r02_gpu_multi_tests.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/r02_bench_p22_n2.json 2> gpurun_out/r02_bench_p22_n2.err
tail -c 2500 gpurun_out/r02_bench_p22_n2.json; tail -5 gpurun_out/r02_bench_p22_n2.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference --ref-power 10) > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err
tail -c 1500 gpurun_out/r02_bench_ref_n2.json
