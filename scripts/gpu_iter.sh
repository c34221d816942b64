# one development iteration on the GPU box: kernel tests, engine parity tests, per-shape profile, graph timings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 --tb=short ${K:+-k "$K"} > gpurun_out/kernel_tests.log 2>&1
echo "== kernel tests exit $?"; tail -n 5 gpurun_out/kernel_tests.log
if [ "${ENGINE:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 600 --tb=short -s > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 8 gpurun_out/engine_tests.log
fi
if [ "${GEMM:-0}" = "1" ]; then
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1
echo "== bench_gemm exit $?"; cat gpurun_out/bench_gemm.txt
fi
if [ "${SHAPES:-1}" = "1" ]; then
timeout 600 python scripts/profile_step.py --batch 8 --shapes --no-profile --out gpurun_out/profile_shapes.txt > gpurun_out/profile_shapes.log 2>&1
echo "== shapes exit $?"
fi
timeout 600 python scripts/profile_step.py --batch 8 --no-profile --out gpurun_out/profile_quick.txt > gpurun_out/profile_quick.log 2>&1
echo "== profile exit $?"; grep -E "graph:|restore" gpurun_out/profile_quick.log | tail -n 3
