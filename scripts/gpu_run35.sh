cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt35.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -6 gpurun_out/pt35.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke35.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke35.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_r35.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_r35.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
