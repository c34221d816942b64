# quick GPU visit: kernel unit tests (optionally filtered by $K), then the step profile
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 --tb=short ${K:+-k "$K"} > gpurun_out/kernel_tests.log 2>&1
echo "== kernel tests exit $?"; tail -n 40 gpurun_out/kernel_tests.log
if [ "${ENGINE:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 600 --tb=short -s > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 15 gpurun_out/engine_tests.log
fi
if [ "${PROFILE:-1}" = "1" ]; then
timeout 600 python scripts/profile_step.py --batch 8 ${PROFILE_ARGS:-} > gpurun_out/profile_step.log 2>&1
echo "== profile exit $?"; tail -n 75 gpurun_out/profile_step.log
fi
