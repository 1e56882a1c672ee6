cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 \"\""
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2p_bench.err | tail -3
CPPF_FRAME_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2p_bench_nopdl.json 2>/dev/null; echo "nopdl rc=$?"
CPPF_TC_SPLIT_NARROW=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2p_bench_nosplit.json 2>/dev/null; echo "nosplit rc=$?"
CPPF_FRAME_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2p_bench_graph.json 2> gpurun_out/r2p_bench_graph.err; echo "graph rc=$?"; grep -v "^W" gpurun_out/r2p_bench_graph.err | tail -3
timeout 200 python tools/heads_profile.py 50000 2700 > gpurun_out/r2p_heads_role_cycles.txt 2>&1; echo "prof rc=$?"
python - <<'PY'
import json
for f in ("r2p_bench", "r2p_bench_nopdl", "r2p_bench_nosplit", "r2p_bench_graph"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        st = d["kernels"]["vote_chain"]["stages_ms"]
        print(f, round(d["ms_per_step"], 4), "heads", round(d["kernels"]["heads"]["ms"], 4), {k: round(v, 4) for k, v in st.items()}, "shot", round(d["kernels"]["shot"]["ms"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
