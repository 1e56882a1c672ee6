cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --sharded-log2 22,24 2> gpurun_out/r2g_bench_g$N.err | grep "^{" > gpurun_out/r2g_bench_g$N.json; echo "bench rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2g_bench_g$N.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/vote_sweep.py --min-log2 16 --max-log2 24 --check-max-log2 24 2> gpurun_out/r2g_vote_sweep_g$N.err | grep "^{" > gpurun_out/r2g_vote_sweep_g$N.jsonl; echo "sweep rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2g_vote_sweep_g$N.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/batched_eval.py --frames 256 2> gpurun_out/r2g_batched_eval_g$N.err | grep "^{" > gpurun_out/r2g_batched_eval_g$N.json; echo "eval rc=$?"
cat gpurun_out/r2g_batched_eval_g$N.json
