# Round 3: A/B of this round's switches on ONE box (box-to-box variation is +-1.5 %): graph medians of the 4-step sample
# and the decode with each switch off, then everything on, then the per-shape profile.  bash scripts/gpu_r3_ab.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r3d}
mkdir -p gpurun_out
run() {  # name, env assignment
  env $2 timeout 300 python scripts/profile_step.py --batch 8 --out gpurun_out/${TAG}_ab_$1.txt > gpurun_out/${TAG}_ab_$1.log 2>&1
  echo "== $1 ($2) exit $?"; grep -E "graph:|restore" gpurun_out/${TAG}_ab_$1.log | tail -n 3
}
run all_on EDTR_NOP=1
run halo_off EDTR_CONV_HALO=0
run xattn_off EDTR_XATTN_SMALL=0
run epign_off EDTR_EPILOGUE_GN=0
run all_on2 EDTR_NOP=1
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_profile_shapes.txt > gpurun_out/${TAG}_profile_shapes.log 2>&1
echo "== shapes exit $?"; grep -A 34 "^--- decode: graph-replay" gpurun_out/${TAG}_profile_shapes.txt
