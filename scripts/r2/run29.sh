cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt29.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt29.log | tail -30
python scripts/r2/stepbench.py --tag "256-bit row accesses" --breakdown > gpurun_out/r2/stepbench29.log 2>&1; cat gpurun_out/r2/stepbench29.log
