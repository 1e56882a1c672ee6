cd $GRAFT_REPO_ROOT
timeout 120 python tools/heads_profile.py 50000 2700 2>&1 | grep -v "^W10" | tail -14
