cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
S3D_MLP_BWD=3 timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt11.log 2>&1
echo "== fused gpu tests (two-set bwd) rc=$?" >> gpurun_out/summary.txt; tail -8 gpurun_out/pt11.log >> gpurun_out/summary.txt
for v in 3 2 1; do
S3D_MLP_BWD=$v timeout 600 python scripts/kbench.py --rays 262144 > gpurun_out/kbench11_$v.log 2>&1; echo "kbench bwd=$v rc=$?" >> gpurun_out/summary.txt
grep "^{\"variant" gpurun_out/kbench11_$v.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
