# bash scripts/gpu_r2_bench.sh <tag>: full GPU test suite, smoke, bench line (with reference_gpu + cpu_baseline), c4 bench
cd $GRAFT_REPO_ROOT
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --tb=short > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 25 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -n 5 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --config c4 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
echo "== bench c4 exit $?"; cat gpurun_out/${TAG}_bench_c4.json; tail -n 5 gpurun_out/${TAG}_bench_c4.err
