cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt6.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -15 gpurun_out/pt6.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused.log >> gpurun_out/summary.txt
timeout 900 python bench.py --rays 65536 --no-cpu-baseline > gpurun_out/bench_fused_64k.log 2>&1; echo "bench fused 64k rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused_64k.log >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --rays 65536 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
