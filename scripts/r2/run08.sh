cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "fused or distill or training or checkpoint or marching or scaler or dynamic" > gpurun_out/r2/pt08.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt08.log | head -40
python scripts/r2/stepbench.py --tag "ray kernel" --breakdown > gpurun_out/r2/stepbench08.log 2>&1; cat gpurun_out/r2/stepbench08.log
