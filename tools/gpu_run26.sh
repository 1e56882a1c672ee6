cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r3d_bench.err | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 70 --csv --log-file gpurun_out/r3d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3d_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 200 python tools/shot_sweep.py > gpurun_out/r3d_shot_sweep.jsonl 2>/dev/null; echo "shot rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 16 --max-log2 22 2>/dev/null | grep "^{" > gpurun_out/r3d_vote_sweep_g1.jsonl; echo "sweep rc=$?"
timeout 200 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r3d_vote_only_g1.jsonl
timeout 200 python tools/example_data.py 2>/dev/null | tail -1 > gpurun_out/r3d_example_data.json
timeout 200 python tools/batched_eval.py --frames 64 2>/dev/null | tail -1 > gpurun_out/r3d_batched_eval_g1.json
ls -la gpurun_out/r3d_*
