cd $GRAFT_REPO_ROOT
timeout 200 python tools/debug_frame.py 2>&1 | grep -v "^frame #\|^W10" | tail -60
