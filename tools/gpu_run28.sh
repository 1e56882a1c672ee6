cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3f_pytest.log; tail -3 gpurun_out/r3f_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err
CPPF_ROT_FAST=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3f_bench_exact.json 2>> gpurun_out/r3f_bench.err
CPPF_B200_LIB=$GRAFT_REPO_ROOT/cppf2_b200/libcppf_exp_rot5.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 "" > gpurun_out/r3f_bench_rot5.json 2>> gpurun_out/r3f_bench.err
timeout 200 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads 2>/dev/null | grep "^{" > gpurun_out/r3f_vote_only_g1.jsonl
CPPF_ROT_FAST=0 timeout 200 python tools/vote_sweep.py --min-log2 22 --max-log2 22 --no-heads --clouds halfcyl 2>/dev/null | grep "^{" > gpurun_out/r3f_vote_only_g1_exact.jsonl
