cd $GRAFT_REPO_ROOT
timeout 200 python tools/debug_frame.py 2>&1 | grep -v "^frame #\|^W10" | grep -c "same as per-instance path: True"
timeout 200 python tools/debug_frame.py 2>&1 | grep -v "^frame #\|^W10" | grep "False\|failed\|Error\|all ok" | head
timeout 300 python -m pytest tests/test_gpu_heads.py -m gpu -x -q -s 2>&1 | grep -E "GATE|passed|failed|Error|error" | cut -c1-200 | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2d_bench.err | tail -3
for r in 1 2 4 16; do CPPF_FRAME_REPLICAS=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2d_bench_rep$r.json 2>/dev/null; echo "rep $r rc=$?"; done
CPPF_FRAME_CALL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2d_bench_noframe.json 2> /dev/null; echo "noframe rc=$?"
