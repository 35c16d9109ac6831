cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for v in 3 2; do
S3D_MLP_BWD=$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_mlp_bwd" -s 2 -c 1 -o gpurun_out/prof_bwd${v}_r12 python bench.py --rays 65536 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bwd$v.log 2>&1
echo "ncu bwd$v rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
