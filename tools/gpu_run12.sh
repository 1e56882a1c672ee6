cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --sharded-log2 22 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2k_bench.err | tail -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --opt --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2k_bench_opt.json 2>/dev/null; echo "opt rc=$?"
