cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- heads tests, chain off"; CPPF_TC_TMEM_CHAIN=0 timeout 300 python -m pytest tests/test_gpu_heads.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2o_bench.err | tail -3
CPPF_TC_TMEM_CHAIN=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2o_bench_nochain.json 2>/dev/null; echo "bench nochain rc=$?"
timeout 200 python tools/heads_profile.py 50000 2700 > gpurun_out/r2o_heads_role_cycles.txt 2>&1; echo "prof rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2o_vote_only_g1.jsonl; echo "sweep rc=$?"
python - <<'PY'
import json
for f in ("r2o_bench", "r2o_bench_nochain"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], d["kernels"]["heads"]["ms"], d["kernels"]["vote_chain"]["stages_ms"], d["kernels"]["shot"]["ms"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
