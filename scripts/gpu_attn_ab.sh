# A/B of the attention kernel variants: unit tests + microbench per variant, then the step profile with the best
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2}; do
  EDTR_ATT_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 200 --tb=short -k attention > gpurun_out/attn_tests_v$v.log 2>&1
  echo "== variant $v tests exit $?"; tail -n 3 gpurun_out/attn_tests_v$v.log
  EDTR_ATT_VARIANT=$v timeout 300 python scripts/bench_attn.py > gpurun_out/bench_attn_v$v.txt 2>&1
  echo "== variant $v bench exit $?"; cat gpurun_out/bench_attn_v$v.txt
done
