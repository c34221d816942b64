# Round 2 iteration visit: GPU tests, bench line, per-shape profile.  Usage: bash scripts/gpu_r2_iter.sh <tag> [pytest -k expr]
cd $GRAFT_REPO_ROOT
TAG=${1:-r2x}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short -k "$KEXPR" > gpurun_out/${TAG}_gpu_tests_k.log 2>&1
  echo "== gpu tests (-k $KEXPR) exit $?"; tail -n 30 gpurun_out/${TAG}_gpu_tests_k.log
fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --tb=short > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 30 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -n 5 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_profile_shapes.txt > gpurun_out/${TAG}_profile_shapes.log 2>&1
echo "== shapes exit $?"; head -n 45 gpurun_out/${TAG}_profile_shapes.txt; grep -A 22 "^--- sample: kernel time" gpurun_out/${TAG}_profile_shapes.txt
timeout 300 python scripts/profile_step.py --out gpurun_out/${TAG}_profile_step.txt > gpurun_out/${TAG}_profile_step.log 2>&1
echo "== step exit $?"; head -n 5 gpurun_out/${TAG}_profile_step.txt
