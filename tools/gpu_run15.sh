cd $GRAFT_REPO_ROOT
for i in 1 2 3; do timeout 300 python tools/batched_eval.py --frames 64 2>/dev/null | tail -1 | cut -c1-200; done
timeout 300 python tools/batched_eval.py --frames 256 2>/dev/null | tail -1 | cut -c1-200
