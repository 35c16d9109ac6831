cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "ffmlp" > gpurun_out/r2/pt07.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt07.log | head -40
