# Round-3 (second half of round 2) record visit: GPU tests, smoke, bench lines (ours, reference-gpu, reference), ncu --set full of the dominant kernel,
# ncu launch list (time + DRAM bytes) of one bench step.  bash scripts/gpu_r3_final.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r03z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --tb=short > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 3 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench exit $?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -n 3 gpurun_out/${TAG}_bench.err
EDTR_REF_BUDGET_S=60 timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> /dev/null
echo "== reference arm exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
NCUARGS="--steps 1 --warmup 3 --sustain-seconds 0 --in-flight 1 --no-cpu-baseline --no-reference-gpu"
EDTR_NCU=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm2_kernel -s 300 -c 3 -o gpurun_out/${TAG}_prof_gemm2 -f python bench.py $NCUARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "== ncu full exit $?"
# the 512 x 512 x 128-channel VAE convolutions (halo mode) are the last gemm2 launches of the step's decode
EDTR_NCU=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm2_kernel -s 1236 -c 12 -o gpurun_out/${TAG}_prof_gemm2_vae -f python bench.py $NCUARGS > gpurun_out/${TAG}_ncu_full_vae.log 2>&1
echo "== ncu full (VAE tail) exit $?"
EDTR_NCU=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $NCUARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/${TAG}_launches.csv
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_gemm2_traffic.json > gpurun_out/${TAG}_launches_summary.txt; head -n 24 gpurun_out/${TAG}_launches_summary.txt
timeout 200 python scripts/profile_step.py --shapes --out gpurun_out/${TAG}_profile_shapes.txt > gpurun_out/${TAG}_profile_shapes.log 2>&1
echo "== shapes exit $?"; head -n 4 gpurun_out/${TAG}_profile_shapes.txt
timeout 120 ncu -i gpurun_out/${TAG}_prof_gemm2.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_gemm2_raw.csv 2>/dev/null; echo "== ncu raw export exit $?"
