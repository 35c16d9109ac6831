cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
RAYS=16384 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wide_forward|k_wide_backward|k_wide_wgrad" -s 12 -c 8 -o gpurun_out/r2/prof_wide_r17 python scripts/config5_bench.py > gpurun_out/r2/ncu_wide17.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2/prof_wide_r17.ncu-rep
