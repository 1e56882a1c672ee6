cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_voting.py tests/test_gpu_sharded.py tests/test_gpu_estimator.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2h_bench.err | tail -3
timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2h_vote_only_g1.jsonl
