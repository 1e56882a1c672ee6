cd $GRAFT_REPO_ROOT
N=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --sharded-log2 22 2> gpurun_out/r3g_bench_g$N.err | grep "^{" > gpurun_out/r3g_bench_g$N.json; echo "bench rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r3g_bench_g$N.err | tail -3
