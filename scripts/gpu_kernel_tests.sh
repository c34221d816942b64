cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
i=0
for k in "gemm" "conv3x3" "attention" "groupnorm or layernorm or softmax or upsample or timestep or sampler or validation"; do
  i=$((i+1))
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" --timeout 180 --tb=short > gpurun_out/t$i.log 2>&1
  echo "== group $i ($k) exit $?"; tail -n 25 gpurun_out/t$i.log
done
