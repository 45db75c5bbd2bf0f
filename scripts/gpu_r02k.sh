cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
bash scripts/gpu_sweep_multi.sh 1 r02
