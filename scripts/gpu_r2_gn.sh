cd $GRAFT_REPO_ROOT
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k "groupnorm" 2>&1 | tail -n 8
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 300 --tb=short -k "tiny or s4 or vae" 2>&1 | tail -n 4
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_shapes.txt > gpurun_out/${TAG}_shapes.log 2>&1
head -n 3 gpurun_out/${TAG}_shapes.txt; grep "groupnorm" gpurun_out/${TAG}_shapes.txt | head -n 24
timeout 300 python scripts/profile_step.py --no-profile --out gpurun_out/${TAG}_step.txt > /dev/null 2>&1; head -n 3 gpurun_out/${TAG}_step.txt
