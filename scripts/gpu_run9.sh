cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "grid or scatter or fused" > gpurun_out/pt9.log 2>&1
echo "== gpu tests rc=$?" >> gpurun_out/summary.txt; tail -8 gpurun_out/pt9.log >> gpurun_out/summary.txt
timeout 900 python scripts/kbench.py --rays 262144 --set 0:1,2 --set 1:1,2 --set 2:1,2 > gpurun_out/kbench9.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
grep "^{" gpurun_out/kbench9.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
