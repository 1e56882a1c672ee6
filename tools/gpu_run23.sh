cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --sharded-log2 22,24 2> gpurun_out/r2u_bench_g$N.err | grep "^{" > gpurun_out/r2u_bench_g$N.json; echo "bench rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2u_bench_g$N.err | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/vote_sweep.py --min-log2 16 --max-log2 24 2> /dev/null | grep "^{" > gpurun_out/r2u_vote_sweep_g$N.jsonl; echo "sweep rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/batched_eval.py --frames 256 2> /dev/null | grep "^{" > gpurun_out/r2u_batched_eval_g$N.json; echo "eval rc=$?"
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
