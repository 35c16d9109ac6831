cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorf.py tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "distillation_with" > gpurun_out/pt21.log 2>&1
echo "rc=$?"; tail -60 gpurun_out/pt21.log
