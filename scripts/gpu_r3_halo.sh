# Round 3 visit: halo-mode convolution A/B (descriptor base-offset variants), cross-attention microbenchmark, bench.
cd $GRAFT_REPO_ROOT
TAG=${1:-r3b}
mkdir -p gpurun_out
for mode in 1 2; do
  EDTR_CONV_HALO=$mode timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --tb=line -k "halo_mode or test_conv3x3" > gpurun_out/${TAG}_halo${mode}_tests.log 2>&1
  echo "== halo mode $mode tests exit $?"; tail -n 12 gpurun_out/${TAG}_halo${mode}_tests.log
done
for sw in 1 0; do
  echo "== cross-attention small kernel EDTR_XATTN_SMALL=$sw"
  EDTR_XATTN_SMALL=$sw timeout 120 python scripts/bench_attn.py 2>&1 | grep "Lk=   77"
done
