cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err
timeout 200 python tools/shot_sweep.py > gpurun_out/r3h_shot_sweep.jsonl 2>/dev/null
timeout 200 python tools/e2e_probe.py > gpurun_out/r3h_e2e_probe.txt 2>&1; tail -6 gpurun_out/r3h_e2e_probe.txt
