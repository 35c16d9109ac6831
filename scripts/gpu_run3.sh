cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt3_parity.log 2>&1
echo "== parity rc=$?" >> gpurun_out/summary.txt; tail -12 gpurun_out/pt3_parity.log >> gpurun_out/summary.txt
for t in test_fused_forward_matches_oracle test_fused_backward_matches_oracle test_fused_scatter_matches_oracle test_fused_adam_tables_matches_torch test_fused_trainer_tracks_autograd_trainer; do
  timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -s -k "$t" --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt3_$t.log 2>&1
  echo "== $t rc=$?" >> gpurun_out/summary.txt; grep -E "rel=|^fused|^plain|passed|failed" gpurun_out/pt3_$t.log | head -12 >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused.log >> gpurun_out/summary.txt
timeout 900 python bench.py --rays 262144 --no-cpu-baseline > gpurun_out/bench_fused_256k.log 2>&1; echo "bench fused 256k rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_fused_256k.log >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
