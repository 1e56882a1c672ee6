cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; echo "bench rc=$?"
