cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/vote_sweep.py --min-log2 22 --max-log2 22 --clouds halfcyl --reps 10 2> gpurun_out/r2j_sweep_g$N.err | grep "^{" > gpurun_out/r2j_sweep_g$N.jsonl; echo "sweep rc=$?"
grep -v "^W10\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2j_sweep_g$N.err | tail -5
