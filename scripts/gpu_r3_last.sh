# Final 1-GPU record of the round: GPU tests, smoke, bench (with the fp32_mode block), CPU reference arm.
cd $GRAFT_REPO_ROOT
TAG=${1:-r03zz}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 400 --tb=short > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 3 gpurun_out/${TAG}_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench exit $?"; cut -c1-200 gpurun_out/${TAG}_bench.json; tail -n 3 gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'gemm frac', d['roofline']['frac'], 'whole', d['roofline']['whole_step']['frac'])
print('fp32_mode', json.dumps(d.get('fp32_mode'))[:900])"
