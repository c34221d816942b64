# NCCL log check on 2 GPUs: bench with EDTR_NCCL_LOG=1 must keep algorithm / channel lines in comm.nccl_log
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --sustain-seconds 0 > gpurun_out/r03q_bench_n2.out 2> gpurun_out/r03q_bench_n2.err
echo "exit $?"; grep "^{" gpurun_out/r03q_bench_n2.out > gpurun_out/r03q_bench_n2.json
python -c "
import json
d=json.load(open('gpurun_out/r03q_bench_n2.json'))
print('value', d['value']); print(json.dumps(d['comm'], indent=1)[:2500])"
ls gpurun_out/nccl_n2.* 2>/dev/null | head -3; env | grep -i nccl | head
