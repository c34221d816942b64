# Round 3: UNet / ControlNet GroupNorm statistics from the producing epilogue — tests, then A/B on one box.
cd $GRAFT_REPO_ROOT
TAG=${1:-r3e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short -k "partials or engine" > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 12 gpurun_out/${TAG}_gpu_tests.log
run() {  # name, env assignment
  env $2 timeout 300 python scripts/profile_step.py --batch 8 --out gpurun_out/${TAG}_ab_$1.txt > gpurun_out/${TAG}_ab_$1.log 2>&1
  echo "== $1 ($2) exit $?"; grep -E "graph:|restore" gpurun_out/${TAG}_ab_$1.log | tail -n 3
}
run unet_on EDTR_NOP=1
run unet_off EDTR_EPILOGUE_GN_UNET=0
run min2048k EDTR_EPILOGUE_GN_MIN_KB=2048
run min12000k EDTR_EPILOGUE_GN_MIN_KB=12000
run unet_on2 EDTR_NOP=1
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_profile_shapes.txt > gpurun_out/${TAG}_profile_shapes.log 2>&1
echo "== shapes exit $?"; grep -E "groupnorm" gpurun_out/${TAG}_profile_shapes.txt | head -30; grep -A 12 "^--- sample: kernel time" gpurun_out/${TAG}_profile_shapes.txt
