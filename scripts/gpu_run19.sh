cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "mappers or proxy or color or teacher or field or scaler or ema" 2>&1 | tail -30
