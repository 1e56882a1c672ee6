cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_estimator.py tests/test_gpu_cloud.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/batched_eval.py --frames 64 2>/dev/null | tail -1 | tee gpurun_out/r2f_batched_eval_g1.json
timeout 300 python tools/example_data.py 2>/dev/null | tail -1 > gpurun_out/r2f_example_data.json; cut -c1-400 gpurun_out/r2f_example_data.json
