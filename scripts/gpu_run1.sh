cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for grp in "near_far or morton or march or composite or inference" "grid" "sh_ or freq" "ffmlp_forward" "ffmlp_backward" "proxy or color or losses" "field or teacher or update_extra or reference_named"; do
  tag=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$grp" --timeout 300 -x --no-header -p no:cacheprovider > gpurun_out/pt_$tag.log 2>&1
  echo "== $grp : rc=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/pt_$tag.log >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
