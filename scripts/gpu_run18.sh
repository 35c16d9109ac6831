cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "scaler or ema or adam or trainer" 2>&1 | tail -30
