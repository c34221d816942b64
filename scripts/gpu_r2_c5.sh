cd $GRAFT_REPO_ROOT
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short -k "c5 or c4_full or verbatim or dropin_pipeline" > gpurun_out/${TAG}_tests.log 2>&1
echo "== tests exit $?"; tail -n 20 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "== bench c5 exit $?"; cat gpurun_out/${TAG}_bench_c5.json; tail -n 5 gpurun_out/${TAG}_bench_c5.err
