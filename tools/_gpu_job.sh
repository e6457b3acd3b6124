set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x -k "verif or merge or ratio or shard or msm or power_pairs" > gpurun_out/r02_gpu_tests_msm_reduce.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_msm_reduce.log
for v in base sqrinl mulinl; do
  if [ $v = base ]; then unset SS_LIB; else export SS_LIB=$PWD/snark-setup_b200/csrc/variants/libss_$v.so; fi
  python bench.py --power 20 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<P
import json
for l in open('gpurun_out/ab_$v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['second_kernel']['avg_launch_ms'], json.dumps(d.get('legs',{}))[:300])
P
done
