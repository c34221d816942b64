# Round 3: cross-attention chunking A/B + in-flight count, on one box.
cd $GRAFT_REPO_ROOT
TAG=${1:-r3h}
mkdir -p gpurun_out
run() {  # name, env assignment
  env $2 timeout 300 python scripts/profile_step.py --batch 8 --out gpurun_out/${TAG}_ab_$1.txt > gpurun_out/${TAG}_ab_$1.log 2>&1
  echo "== $1 ($2) exit $?"; grep -E "graph:" gpurun_out/${TAG}_ab_$1.log | head -n 1
}
run auto EDTR_NOP=1
run chunks1 EDTR_XATTN_CHUNKS=1
run chunks2 EDTR_XATTN_CHUNKS=2
run chunks8 EDTR_XATTN_CHUNKS=8
run auto2 EDTR_NOP=1
for nf in 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --in-flight $nf --no-cpu-baseline --no-reference-gpu --no-fp32 --sustain-seconds 0 > gpurun_out/${TAG}_bench_if$nf.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_if$nf.json'))
print('in_flight $nf: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sequential', round(d['sequential']['value'],2))"
done
