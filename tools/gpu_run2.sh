set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_voting.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --sharded-log2 22,24 > gpurun_out/r2b_bench_g2.json 2> gpurun_out/r2b_bench_g2.err; echo "bench rc=$?"
grep -v "^W1017\|^\[W" gpurun_out/r2b_bench_g2.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/vote_sweep.py --min-log2 16 --max-log2 22 > gpurun_out/r2b_vote_sweep_g2.jsonl 2> gpurun_out/r2b_vote_sweep_g2.err; echo "sweep rc=$?"
grep -v "^W1017\|^\[W" gpurun_out/r2b_vote_sweep_g2.err | tail -5
