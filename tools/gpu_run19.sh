cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r2q_bench.err | tail -3
CPPF_VOTE_LANES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2q_bench_nolanes.json 2>/dev/null; echo "nolanes rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2q_vote_only_g1.jsonl; echo "sweep rc=$?"
CPPF_VOTE_LANES=0 timeout 300 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r2q_vote_only_g1_nolanes.jsonl; echo "sweep rc=$?"
python - <<'PY'
import json
for f in ("r2q_bench", "r2q_bench_nolanes"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        st = d["kernels"]["vote_chain"]["stages_ms"]
        print(f, round(d["ms_per_step"], 4), "heads", round(d["kernels"]["heads"]["ms"], 4), {k: round(v, 4) for k, v in st.items()}, "shot", round(d["kernels"]["shot"]["ms"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "ERR", e)
for f in ("r2q_vote_only_g1", "r2q_vote_only_g1_nolanes"):
    for l in open(f"gpurun_out/{f}.jsonl"):
        d = json.loads(l); print(f, d["cloud"], round(d["ms"], 3), round(d["tuples_per_sec"] / 1e6), d["stages_ms_rank0"], d["parity"])
PY
