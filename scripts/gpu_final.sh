# Reduced final visit: GPU tests, bench, ncu --set full of the top kernel, then the ncu launch list (longest, last)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 --tb=short > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests exit $?"; tail -n 3 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
EDTR_NCU=1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm2_kernel -s 300 -c 3 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $?"
EDTR_NCU=1 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/launches.csv
python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm2_traffic.json > gpurun_out/launches_summary.txt; head -n 12 gpurun_out/launches_summary.txt
