cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2t_bench.err | tail -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2t_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_tc|frame_vote_center|frame_rotation|frame_shot_descriptor|frame_shot_normals|frame_select|frame_fold" -s 40 -c 10 -o gpurun_out/r2t_frame python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2t_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 400 python tools/vote_sweep.py --min-log2 16 --max-log2 24 --cpu-max-log2 18 2>/dev/null | grep "^{" > gpurun_out/r2t_vote_sweep_g1.jsonl; echo "sweep rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2t_vote_only_g1.jsonl
timeout 300 python tools/example_data.py 2>/dev/null | tail -1 > gpurun_out/r2t_example_data.json
timeout 300 python tools/batched_eval.py --frames 64 2>/dev/null | tail -1 > gpurun_out/r2t_batched_eval_g1.json
ls -la gpurun_out/r2t_*
