cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 300 --no-header -p no:cacheprovider > gpurun_out/r2/pt12.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt12.log | head -30
L=gpurun_out/r2/stepbench12.log; : > $L
S3D_FUSED_BACKWARD=0 timeout 300 python scripts/r2/stepbench.py --tag "separate mlp bwd + scatter" --breakdown >> $L 2>&1
timeout 300 python scripts/r2/stepbench.py --tag "fused mlp bwd + scatter" --breakdown >> $L 2>&1
cat $L
