# ncu --set full captures of single conv / gemm launches (dev tool); reports land in gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {  # name kernel-regex env kind args...
  name=$1; shift; kre=$1; shift; envs=$1; shift
  env $envs timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f \
      -o gpurun_out/ncu_$name python scripts/ncu_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "== $name exit $?"; tail -n 2 gpurun_out/ncu_$name.log
}
run conv320_v2 gemm2_kernel X=1 conv 8 64 64 320 320
run conv320_v1 gemm_conv_kernel EDTR_GEMM_V1=1 conv 8 64 64 320 320
run conv1280_v2 gemm2_kernel X=1 conv 8 16 16 1280 1280
run conv256_v2 gemm2_kernel X=1 conv 8 256 256 256 256
run conv128_v2 gemm2_kernel X=1 conv 8 512 512 128 128
run gemm320_v2 gemm2_kernel X=1 gemm 32768 320 320
ls -la gpurun_out/*.ncu-rep
