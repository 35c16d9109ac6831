cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt20.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -8 gpurun_out/pt20.log >> gpurun_out/summary.txt
timeout 600 python scripts/kbench.py --set 10:-1,50000,120000,200000,300000,800000,2000000 --set 11:-1,100000,250000,600000,1500000,4000000 > gpurun_out/kbench20.log 2>&1
echo "kbench rc=$?" >> gpurun_out/summary.txt
grep -v "^\[" gpurun_out/kbench20.log | tail -20 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
