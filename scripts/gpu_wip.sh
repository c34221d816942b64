# GPU check of the unverified branches: attention lazy-max variant, SwinIR kernels + drop-in.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
if [ -d .wip/attn ]; then
  cd .wip/attn
  for v in 0 1; do
    EDTR_ATT_LAZYMAX=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k attention > $GRAFT_REPO_ROOT/gpurun_out/attn_lazy$v.log 2>&1
    echo "== EDTR_ATT_LAZYMAX=$v attention tests exit $?"; tail -n 3 $GRAFT_REPO_ROOT/gpurun_out/attn_lazy$v.log
    EDTR_ATT_LAZYMAX=$v timeout 120 python scripts/bench_attn.py 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/bench_attn_lazy$v.txt | head -n 4
  done
  EDTR_ATT_K128=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 --tb=short -k attention > $GRAFT_REPO_ROOT/gpurun_out/attn_k128.log 2>&1
  echo "== EDTR_ATT_K128=1 attention tests exit $?"; tail -n 3 $GRAFT_REPO_ROOT/gpurun_out/attn_k128.log
  EDTR_ATT_K128=1 timeout 120 python scripts/bench_attn.py 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/bench_attn_k128.txt | head -n 4
  cd $GRAFT_REPO_ROOT
fi
if [ -d .wip/swinir ]; then
  cd .wip/swinir
  timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 300 --tb=short \
      -k "swinir or window_attention or layernorm_padded or pixel_unshuffle or pointwise_activations or leaky_relu" > $GRAFT_REPO_ROOT/gpurun_out/swinir_tests.log 2>&1
  echo "== swinir tests exit $?"; tail -n 25 $GRAFT_REPO_ROOT/gpurun_out/swinir_tests.log
  cd $GRAFT_REPO_ROOT
fi
