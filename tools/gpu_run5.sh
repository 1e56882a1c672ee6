cd $GRAFT_REPO_ROOT
timeout 200 python tools/debug_frame.py 2>&1 | grep -v "^frame #\|^W10" | grep -c "same as per-instance path: True"
timeout 200 python tools/debug_frame.py 2>&1 | grep -v "^frame #\|^W10" | grep "False\|failed\|Error\|all ok" | head
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2e_bench.err | tail -3
CPPF_FRAME_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2e_bench_graph.json 2>/dev/null; echo "graph rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
