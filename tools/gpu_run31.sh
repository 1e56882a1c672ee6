cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err; echo "bench rc=$?"; grep -v "^W" gpurun_out/r3i_bench.err | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 70 --csv --log-file gpurun_out/r3i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3i_ncu.log 2>&1; echo "ncu launches rc=$?"
