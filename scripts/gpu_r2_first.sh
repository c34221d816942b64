# Round 2, first GPU visit: state of main after the swinir merge, the attn-wip variants, per-shape kernel times,
# and the reference GPU arm (stock PyTorch kernels) on the same box.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --tb=short > gpurun_out/r2a_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 15 gpurun_out/r2a_gpu_tests.log
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 3 > gpurun_out/r2a_bench_refgpu.json 2> gpurun_out/r2a_bench_refgpu.err
echo "== reference-gpu exit $?"; cat gpurun_out/r2a_bench_refgpu.json; tail -n 5 gpurun_out/r2a_bench_refgpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "== bench exit $?"; cat gpurun_out/r2a_bench.json; tail -n 5 gpurun_out/r2a_bench.err
timeout 300 python scripts/profile_step.py --shapes --out gpurun_out/r2a_profile_shapes.txt > gpurun_out/r2a_profile_shapes.log 2>&1
echo "== shapes exit $?"; head -n 70 gpurun_out/r2a_profile_shapes.txt
timeout 200 python scripts/bench_gemm.py > gpurun_out/r2a_bench_gemm.txt 2>&1
echo "== bench_gemm exit $?"; cat gpurun_out/r2a_bench_gemm.txt
if [ -d .wip/attn ]; then
  cd .wip/attn
  for v in "EDTR_ATT_LAZYMAX=0" "EDTR_ATT_LAZYMAX=1" "EDTR_ATT_K128=1"; do
    env $v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k attention > $GRAFT_REPO_ROOT/gpurun_out/r2a_attn_$v.log 2>&1
    echo "== $v attention tests exit $?"; tail -n 3 $GRAFT_REPO_ROOT/gpurun_out/r2a_attn_$v.log
    env $v timeout 120 python scripts/bench_attn.py 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/r2a_bench_attn_$v.txt | head -n 12
  done
  cd $GRAFT_REPO_ROOT
fi
