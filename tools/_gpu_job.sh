set -x
cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_final_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | tail -3
tail -c 200 gpurun_out/r02_bench_final_n$N.json; grep "bench " gpurun_out/bench_n$N.err | tail -5
