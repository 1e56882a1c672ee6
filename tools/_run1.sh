set -x
cd /root/repo
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3a_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
CPPF_SELECT_CLUSTER=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3a_bench_noclu.json 2>> gpurun_out/r3a_bench.err
timeout 200 python tools/shot_sweep.py > gpurun_out/r3a_shot_sweep.jsonl 2>> gpurun_out/r3a_bench.err
tail -5 gpurun_out/r3a_pytest.log
