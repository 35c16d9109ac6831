cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_tensorf.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/r2/pt19.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt19.log | head -30
timeout 600 python scripts/config5_bench.py > gpurun_out/r2/config5_19.log 2>&1; tail -1 gpurun_out/r2/config5_19.log
timeout 300 python scripts/r2/wide_micro.py > gpurun_out/r2/wide_micro19.log 2>&1; cat gpurun_out/r2/wide_micro19.log
