# quick A/B of the 4-step sample / decode graphs: bash scripts/gpu_r2_quick.sh <tag> [ENV=VAL ...]
cd $GRAFT_REPO_ROOT
TAG=${1:-q}; shift
mkdir -p gpurun_out
env "$@" timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_profile_shapes.txt > gpurun_out/${TAG}_profile_shapes.log 2>&1
echo "== shapes exit $?"; head -n 40 gpurun_out/${TAG}_profile_shapes.txt
env "$@" timeout 300 python scripts/profile_step.py --no-profile --out gpurun_out/${TAG}_profile_step.txt > gpurun_out/${TAG}_profile_step.log 2>&1
echo "== step exit $?"; head -n 5 gpurun_out/${TAG}_profile_step.txt
