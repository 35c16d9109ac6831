cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt
for n in 8 4 2; do
NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_n$n.log 2>&1; echo "bench n$n rc=$?" >> gpurun_out/summary.txt
grep "^{" gpurun_out/bench_n$n.log | tail -1 >> gpurun_out/summary.txt
grep -m3 -i "nvls\|Connected all trees\|Channel 00" gpurun_out/bench_n$n.log >> gpurun_out/summary.txt
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_n1.log >> gpurun_out/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/dp_consistency.py > gpurun_out/dp_consistency8.log 2>&1; echo "dp consistency rc=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/dp_consistency8.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
