# GroupNorm 16-CTA cluster A/B, ncu capture of the tile-split attention kernel, per-shape profile
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k "groupnorm or attention" > gpurun_out/gn_tests.log 2>&1
echo "== groupnorm/attention tests exit $?"; tail -n 5 gpurun_out/gn_tests.log
for c in 0 1; do
  EDTR_GN_CS16=$c timeout 300 python scripts/profile_step.py --batch 8 --out gpurun_out/profile_cs16_$c.txt > gpurun_out/profile_cs16_$c.log 2>&1
  echo "== EDTR_GN_CS16=$c profile exit $?"; grep -E "graph:|restore" gpurun_out/profile_cs16_$c.log | tail -n 3; grep -E "groupnorm" gpurun_out/profile_cs16_$c.log | head -n 8
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_ts_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_attn_ts python scripts/ncu_attn.py > gpurun_out/ncu_attn_ts.log 2>&1
echo "== ncu attention exit $?"
timeout 600 python scripts/profile_step.py --batch 8 --shapes --no-profile --out gpurun_out/profile_shapes.txt > gpurun_out/profile_shapes.log 2>&1
echo "== shapes exit $?"
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 400 --tb=short > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 4 gpurun_out/engine_tests.log
