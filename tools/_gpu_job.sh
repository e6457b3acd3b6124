set -x
cd $GRAFT_REPO_ROOT
export SS_CONCURRENT_VECTORS=0
python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/plain_power18_final.json 2>/dev/null; echo "plain exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_ncu_launches_power18_final.csv python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2> gpurun_out/ncu.err; echo "ncu exit $?"
