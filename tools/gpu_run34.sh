cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
