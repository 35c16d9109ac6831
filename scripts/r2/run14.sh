cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt14.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt14.log | tail -20
python scripts/r2/stepbench.py --tag "thread-per-sample write pass" --breakdown > gpurun_out/r2/stepbench14.log 2>&1; cat gpurun_out/r2/stepbench14.log
timeout 600 python scripts/config5_bench.py > gpurun_out/r2/config5_14.log 2>&1; tail -2 gpurun_out/r2/config5_14.log
