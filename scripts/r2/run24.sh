cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_tensorf.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt24.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt24.log | tail -30
timeout 600 python scripts/config5_bench.py > gpurun_out/r2/config5_24.log 2>&1; tail -1 gpurun_out/r2/config5_24.log
