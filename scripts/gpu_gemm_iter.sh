# GEMM kernel iteration: kernel unit tests (gemm/conv), per-shape bench, whole-graph timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 --tb=short > gpurun_out/kernel_tests.log 2>&1
echo "== kernel tests exit $?"; tail -n 5 gpurun_out/kernel_tests.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1
echo "== bench_gemm exit $?"; cat gpurun_out/bench_gemm.txt
timeout 600 python scripts/profile_step.py --batch 8 --no-profile --out gpurun_out/profile_quick.txt > gpurun_out/profile_quick.log 2>&1
echo "== profile exit $?"; grep -E "graph:|restore" gpurun_out/profile_quick.log | tail -n 3
