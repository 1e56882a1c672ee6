set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python tools/shot_sweep.py > gpurun_out/r2a_shot_sweep.jsonl 2>&1; echo "shot rc=$?"
timeout 300 python bench.py --impl reference --shot-sweep > gpurun_out/r2a_shot_sweep_cpu.jsonl 2>&1; echo "shotcpu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vote_center_kernel|rotation_hist_kernel|shot_descriptor_kernel|shot_normals_kernel" -s 108 -c 36 -o gpurun_out/r2a_vote_shot env CPPF_STREAMS=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/vote_sweep.py --min-log2 16 --max-log2 22 --cpu-max-log2 18 > gpurun_out/r2a_vote_sweep_g1.jsonl 2> gpurun_out/r2a_vote_sweep_g1.err; echo "sweep rc=$?"
ls -la gpurun_out | tail -20
