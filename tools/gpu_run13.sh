cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 10 --warmup 3 --sharded-log2 "" > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2l_bench.err | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2l_bench20.json 2>/dev/null; echo "bench20 rc=$?"
