set -x
cd $GRAFT_REPO_ROOT
python -X faulthandler bench.py > gpurun_out/dbg_bench.json 2> gpurun_out/dbg_bench.err; echo "exit $?"
tail -60 gpurun_out/dbg_bench.err; tail -c 300 gpurun_out/dbg_bench.json
