cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt05.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2/pt05.log | tail -30
L=gpurun_out/r2/stepbench05.log; : > $L
S3D_MLP_FWD=ss python scripts/r2/stepbench.py --tag "fwd=ss" --breakdown >> $L 2>&1
python scripts/r2/stepbench.py --tag "fwd=ts ctas=3" --breakdown >> $L 2>&1
S3D_MLP_FWD_CTAS=2 python scripts/r2/stepbench.py --tag "fwd=ts ctas=2" --breakdown >> $L 2>&1
S3D_MLP_FWD_CTAS=4 python scripts/r2/stepbench.py --tag "fwd=ts ctas=4" --breakdown >> $L 2>&1
cat $L
