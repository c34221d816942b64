# bash scripts/gpu_r2_multi.sh <tag> <ngpus>: multi-GPU tests + bench lines (c2/c3 image-parallel and c4 tile-parallel)
cd $GRAFT_REPO_ROOT
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 600 --tb=short -k "nccl or non_current_device" > gpurun_out/${TAG}_multi_tests.log 2>&1
echo "== multi-GPU tests exit $?"; tail -n 25 gpurun_out/${TAG}_multi_tests.log
EDTR_NCCL_LOG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.out 2> gpurun_out/${TAG}_bench_n$N.err
grep "^{" gpurun_out/${TAG}_bench_n$N.out > gpurun_out/${TAG}_bench_n$N.json; echo "NCCL lines on stdout: $(grep -c NCCL gpurun_out/${TAG}_bench_n$N.out), on stderr: $(grep -c NCCL gpurun_out/${TAG}_bench_n$N.err)"; grep -h -E "NCCL INFO.*(NVLS|Connected all|nranks|Channel 00/)" gpurun_out/${TAG}_bench_n$N.out gpurun_out/${TAG}_bench_n$N.err | head -n 8
echo "== bench N=$N exit $?"; cat gpurun_out/${TAG}_bench_n$N.json | cut -c1-600; python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'comm',json.dumps(d.get('comm'))[:1500])"; tail -n 5 gpurun_out/${TAG}_bench_n$N.err
EDTR_NCCL_LOG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --config c4 --steps 3 --warmup 3 | grep "^{" > gpurun_out/${TAG}_bench_c4_n$N.json 2> gpurun_out/${TAG}_bench_c4_n$N.err
echo "== bench c4 N=$N exit $?"; cat gpurun_out/${TAG}_bench_c4_n$N.json; tail -n 5 gpurun_out/${TAG}_bench_c4_n$N.err
ls gpurun_out/nccl_* 2>/dev/null | head; for f in gpurun_out/nccl_n$N.*; do grep -h -E 'NVLS|Connected all|AllGather|AllReduce|nranks|Channel 00' $f | head -n 12; break; done
