set -x
cd $GRAFT_REPO_ROOT
N=${NGPU:-4}
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3) > gpurun_out/r02_bench_p22_n$N.json 2> gpurun_out/r02_bench_p22_n$N.err
tail -3 gpurun_out/r02_bench_p22_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_p22_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['legs'], round(d['e2e']['value']), round(d['e2e']['pageable']['value']), d['gpu_launches'], d['verdict_all_steps'], d['parity_spot_check'], d['clocks'])
P
