cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 600 --tb=short -s > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 40 gpurun_out/engine_tests.log
timeout 600 python scripts/profile_step.py --batch 8 > gpurun_out/profile_step.log 2>&1
echo "== profile exit $?"; tail -n 70 gpurun_out/profile_step.log
