# Round 3: fp32 mode — kernel tests, engine tests, timing of one fp32 restore (B=1 and B=8).
cd $GRAFT_REPO_ROOT
TAG=${1:-r3f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 400 --tb=short -k "f32 or fp32" > gpurun_out/${TAG}_fp32_tests.log 2>&1
echo "== fp32 tests exit $?"; tail -n 25 gpurun_out/${TAG}_fp32_tests.log
timeout 600 python scripts/bench_fp32.py > gpurun_out/${TAG}_fp32_timing.txt 2>&1
echo "== fp32 timing exit $?"; tail -n 12 gpurun_out/${TAG}_fp32_timing.txt
