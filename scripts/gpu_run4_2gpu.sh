cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -s --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt4_fused.log 2>&1
echo "== fused tests rc=$?" >> gpurun_out/summary.txt; grep -E "emulation|passed|failed" gpurun_out/pt4_fused.log | head -12 >> gpurun_out/summary.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_n1.log >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_n2.log >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_consistency.py > gpurun_out/dp_consistency.log 2>&1; echo "dp consistency rc=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/dp_consistency.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
