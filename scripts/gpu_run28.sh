cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "reference_outputs" > gpurun_out/pt28.log 2>&1
echo "rc=$?"; tail -40 gpurun_out/pt28.log
