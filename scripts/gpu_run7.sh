cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt7.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -15 gpurun_out/pt7.log >> gpurun_out/summary.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r7.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_r7.log >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_|k_march" -s 20 -c 9 -o gpurun_out/prof_field_r7 python bench.py --rays 65536 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full7.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_grid_forward" -c 4 -o gpurun_out/prof_grid_r7 python bench.py --rays 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_grid7.log 2>&1
echo "ncu grid rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
