cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt10.log 2>&1
echo "== fused gpu tests (two-chain bwd) rc=$?" >> gpurun_out/summary.txt; tail -8 gpurun_out/pt10.log >> gpurun_out/summary.txt
S3D_MLP_BWD=1 timeout 600 python scripts/kbench.py --rays 262144 > gpurun_out/kbench10_serial.log 2>&1; echo "kbench serial rc=$?" >> gpurun_out/summary.txt
grep "^{" gpurun_out/kbench10_serial.log >> gpurun_out/summary.txt
S3D_MLP_BWD=2 timeout 600 python scripts/kbench.py --rays 262144 > gpurun_out/kbench10_two.log 2>&1; echo "kbench two-chain rc=$?" >> gpurun_out/summary.txt
grep "^{" gpurun_out/kbench10_two.log >> gpurun_out/summary.txt
S3D_MLP_BWD=2 timeout 600 python scripts/kbench.py --rays 65536 > gpurun_out/kbench10_two64.log 2>&1; echo "kbench two-chain 64k rc=$?" >> gpurun_out/summary.txt
grep "^{" gpurun_out/kbench10_two64.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
