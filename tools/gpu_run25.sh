cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3c_pytest.log; tail -3 gpurun_out/r3c_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
CPPF_VOTE_OCC=6 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3c_bench_occ6.json 2>> gpurun_out/r3c_bench.err
timeout 200 python tools/shot_sweep.py > gpurun_out/r3c_shot_sweep.jsonl 2>> gpurun_out/r3c_bench.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/r3c_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_smoke.py > gpurun_out/r3c_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r3c_memcheck.log; grep -c "Race reported\|hazard" gpurun_out/r3c_racecheck.log; tail -3 gpurun_out/r3c_racecheck.log
