cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for grp in "march or grid or field or teacher or update_extra or reference_named"; do
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$grp" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt2_parity.log 2>&1
  echo "== parity subset rc=$?" >> gpurun_out/summary.txt; tail -8 gpurun_out/pt2_parity.log >> gpurun_out/summary.txt
done
for t in test_fused_forward_matches_oracle test_fused_backward_matches_oracle test_fused_adam_tables_matches_torch test_fused_trainer_tracks_autograd_trainer; do
  timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "$t" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt2_$t.log 2>&1
  echo "== $t rc=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pt2_$t.log >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 5 --rays 262144 --no-cpu-baseline > gpurun_out/bench_fused_256k.log 2>&1; echo "bench fused 256k rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused_256k.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 5 --engine autograd --no-cpu-baseline > gpurun_out/bench_autograd.log 2>&1; echo "bench autograd rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_autograd.log >> gpurun_out/summary.txt
# profiles: launch list of a short fused bench, then full captures of the two gather/scatter kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_encode|k_ngp_scatter|k_ngp_mlp" -s 12 -c 8 -o gpurun_out/prof_field_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
