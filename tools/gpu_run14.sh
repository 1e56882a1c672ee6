cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2m_bench.err | tail -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 16 --max-log2 24 --cpu-max-log2 18 2>/dev/null | grep "^{" > gpurun_out/r2m_vote_sweep_g1.jsonl; echo "sweep rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2m_vote_only_g1.jsonl
timeout 300 python tools/shot_sweep.py > gpurun_out/r2m_shot_sweep.jsonl 2>/dev/null; echo "shot rc=$?"
timeout 300 python tools/example_data.py 2>/dev/null | tail -1 > gpurun_out/r2m_example_data.json
timeout 300 python tools/batched_eval.py --frames 64 2>/dev/null | tail -1 > gpurun_out/r2m_batched_eval_g1.json
