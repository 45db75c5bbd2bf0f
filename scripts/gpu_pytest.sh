# usage: bash scripts/gpu_pytest.sh <pytest args...>   (on the GPU box; hang protection)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest -q -m gpu -x "$@" 2>&1 | tail -25
