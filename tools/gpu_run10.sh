cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_voting.py tests/test_gpu_estimator.py -m gpu -x -q 2>&1 | tail -2
for r in 1 2 4 8; do CPPF_FRAME_REPLICAS=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2i_bench_rep$r.json 2>/dev/null; echo "rep $r rc=$?"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_tc|frame_vote_center|frame_rotation|frame_shot_descriptor|frame_shot_normals|frame_select|frame_fold" -s 40 -c 10 -o gpurun_out/r2i_frame python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
