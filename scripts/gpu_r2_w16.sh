cd $GRAFT_REPO_ROOT
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short 2>&1 | tail -n 12
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -n 6
timeout 300 python scripts/bench_epilogue.py 2>&1 | tee gpurun_out/${TAG}_bench_epilogue.txt
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_shapes.txt > gpurun_out/${TAG}_shapes.log 2>&1
head -n 45 gpurun_out/${TAG}_shapes.txt
timeout 300 python scripts/profile_step.py --no-profile --out gpurun_out/${TAG}_step.txt > /dev/null 2>&1; head -n 3 gpurun_out/${TAG}_step.txt
