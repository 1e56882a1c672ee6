cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3b_pytest.log; tail -3 gpurun_out/r3b_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err
timeout 200 python tools/shot_sweep.py > gpurun_out/r3b_shot_sweep.jsonl 2>> gpurun_out/r3b_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 70 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sharded-log2 '' > gpurun_out/r3b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:frame_shot|frame_select|frame_backvote|frame_decode|frame_fold|frame_prep|frame_pose' -s 40 -c 24 -o gpurun_out/r3b_frame_small python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sharded-log2 '' > gpurun_out/r3b_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
