set -x
cd $GRAFT_REPO_ROOT
run() { name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ab_tile_${name}.json 2> gpurun_out/r02_ab_tile_${name}.err
}
run base SS_DUMMY=1
run tile2x SS_TILE_ELEMS=1212416
run ratio21 SS_RATIO_TILE_LOG2=21
run both SS_TILE_ELEMS=1212416 SS_RATIO_TILE_LOG2=21
python - <<'P'
import json
for name in ('base','tile2x','ratio21','both'):
    try:
      d=json.loads(open(f'gpurun_out/r02_ab_tile_{name}.json').read().strip().splitlines()[-1])
      print(name, round(d['value']), round(d['legs']['contribute']['ms_per_step'],1), round(d['legs']['verify']['ms_per_step'],1), d['verdict_all_steps'], d['parity_spot_check'])
    except Exception as e:
      print(name,'ERR',e); print(open(f'gpurun_out/r02_ab_tile_{name}.err').read()[-800:])
P
