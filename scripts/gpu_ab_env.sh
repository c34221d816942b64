# A/B of one environment switch on the step profile: AB_VAR=<name> [AB_VALUES="0 1"], after the kernel + engine tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 200 --tb=short > gpurun_out/kernel_tests.log 2>&1
echo "== kernel tests exit $?"; tail -n 5 gpurun_out/kernel_tests.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 600 --tb=short > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 5 gpurun_out/engine_tests.log
for v in ${AB_VALUES:-0 1}; do
  env ${AB_VAR}=$v timeout 300 python scripts/profile_step.py --batch 8 --out gpurun_out/profile_${AB_VAR}_$v.txt > gpurun_out/profile_${AB_VAR}_$v.log 2>&1
  echo "== ${AB_VAR}=$v profile exit $?"; grep -E "graph:|restore" gpurun_out/profile_${AB_VAR}_$v.log | tail -n 3
  grep -E "gemm_conv_kernel|gemm2_kernel" gpurun_out/profile_${AB_VAR}_$v.log | head -n 10
done
