cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --sharded-log2 22 2> gpurun_out/r3e_bench_g$N.err | grep "^{" > gpurun_out/r3e_bench_g$N.json; echo "bench rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r3e_bench_g$N.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/vote_sweep.py --min-log2 16 --max-log2 22 2> /dev/null | grep "^{" > gpurun_out/r3e_vote_sweep_g$N.jsonl; echo "sweep rc=$?"
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/shot_sweep.py > gpurun_out/r3e_shot_sweep.jsonl 2>/dev/null; echo "shot rc=$?"
