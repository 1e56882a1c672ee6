set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_estimator.py -m gpu -x -q 2>&1 | tail -25
test ${PIPESTATUS[0]} -eq 0 || exit 1
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "GATE|passed|failed|Error|error" | tail -30
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2c_bench.err | tail -3
CPPF_FRAME_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2c_bench_graph.json 2> gpurun_out/r2c_bench_graph.err; echo "bench graph rc=$?"; grep -v "^W" gpurun_out/r2c_bench_graph.err | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/example_data.py > gpurun_out/r2c_example_data.json 2> gpurun_out/r2c_example_data.err; echo "example rc=$?"; tail -2 gpurun_out/r2c_example_data.err
