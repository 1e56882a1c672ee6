cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --sharded-log2 22,24 2> gpurun_out/r2n_bench_g$N.err | grep "^{" > gpurun_out/r2n_bench_g$N.json; echo "bench rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2n_bench_g$N.err | tail -3
CPPF_FRAME_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 --sharded-log2 "" 2> gpurun_out/r2n_bench_graph_g$N.err | grep "^{" > gpurun_out/r2n_bench_graph_g$N.json; echo "bench graph rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/batched_eval.py --frames 256 2> /dev/null | grep "^{" > gpurun_out/r2n_batched_eval_g$N.json; echo "eval rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/batched_eval.py --frames 1024 2> /dev/null | grep "^{" > gpurun_out/r2n_batched_eval1024_g$N.json; echo "eval rc=$?"
