cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt13.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -6 gpurun_out/pt13.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke13.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke13.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_r13.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_r13.log >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r13.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_ref_r13.log >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r13.csv python bench.py --rays 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch13.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
