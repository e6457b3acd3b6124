set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x -k "verif or merge or ratio or shard or msm or power_pairs or mnt" 2>&1 | tail -2
python tools/shard_probe.py 22 8 1
python tools/shard_probe.py 22 16 1
