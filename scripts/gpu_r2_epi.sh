cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k "gemm or conv" 2>&1 | tail -n 3
timeout 300 python scripts/bench_epilogue.py 2>&1 | tee gpurun_out/${TAG}_bench_epilogue.txt
