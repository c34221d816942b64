# per-shape kernel time of one graph replay (single stream), then the PDL on/off comparison of the whole graphs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py --batch 8 --shapes --no-profile --out gpurun_out/profile_shapes.txt > gpurun_out/profile_shapes.log 2>&1
echo "== shapes exit $?"; tail -n 3 gpurun_out/profile_shapes.log
for pdl in 0 1; do
EDTR_PDL=$pdl timeout 600 python scripts/profile_step.py --batch 8 --no-profile --out gpurun_out/profile_pdl$pdl.txt > gpurun_out/profile_pdl$pdl.log 2>&1
echo "== pdl=$pdl exit $?"; grep -E "graph:|restore" gpurun_out/profile_pdl$pdl.log | tail -n 3
done
