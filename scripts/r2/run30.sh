cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 120 scripts/r2/bin/red_micro | tee gpurun_out/r2/red_micro30.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "ffmlp" > gpurun_out/r2/pt30.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt30.log | tail -10
