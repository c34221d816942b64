# One GPU-box visit: unit tests, engine parity, bench, ncu launch list, one full ncu capture of the top kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --tb=short -s > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 30 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
EDTR_NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/launches.csv
python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm2_traffic.json > gpurun_out/launches_summary.txt; head -n 12 gpurun_out/launches_summary.txt
EDTR_NCU=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm2_kernel -s 300 -c 3 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_ts_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_attn python scripts/ncu_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "== ncu attention exit $?"; ls -la gpurun_out/*.ncu-rep
fi
timeout 600 python scripts/profile_step.py --batch 8 > gpurun_out/profile_step.log 2>&1
echo "== profile exit $?"; tail -n 60 gpurun_out/profile_step.log
