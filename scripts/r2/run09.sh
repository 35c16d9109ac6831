cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt09.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2/pt09.log | tail -30
python scripts/r2/stepbench.py --tag "bbox-clipped marcher + merged ray scans" --breakdown > gpurun_out/r2/stepbench09.log 2>&1; cat gpurun_out/r2/stepbench09.log
