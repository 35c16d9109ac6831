cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt06.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2/pt06.log | tail -30
timeout 900 python scripts/ref_ab.py > gpurun_out/r2/ref_ab06.log 2>&1; echo "ref_ab rc=$?"; tail -25 gpurun_out/r2/ref_ab06.log
timeout 600 python scripts/config23_bench.py > gpurun_out/r2/config23_06.log 2>&1; echo "config23 rc=$?"; tail -12 gpurun_out/r2/config23_06.log
