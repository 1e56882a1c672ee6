cd $GRAFT_REPO_ROOT
for e in base 64 128 256; do
  echo "=== exp $e"
  if [ $e = base ]; then timeout 200 python tools/heads_profile.py 50000 2700 2>&1 | grep -v "^W" | tail -12
  else CPPF_B200_LIB=$GRAFT_REPO_ROOT/cppf2_b200/libcppf_exp$e.so timeout 200 python tools/heads_profile.py 50000 2700 2>&1 | grep -v "^W" | tail -12; fi
done > gpurun_out/r2r_heads_experiments.txt 2>&1
cat gpurun_out/r2r_heads_experiments.txt | grep "===\|forward\|epilogue s0\|issuer s0"
