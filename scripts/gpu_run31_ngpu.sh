cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
n=${NGPU:-2}
rm -f gpurun_out/summary_n$n.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench31_n$n.log 2>&1; echo "bench n$n rc=$?" >> gpurun_out/summary_n$n.txt
grep "^{" gpurun_out/bench31_n$n.log | tail -1 >> gpurun_out/summary_n$n.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 scripts/dp_consistency.py > gpurun_out/dp_consistency31_n$n.log 2>&1; echo "dp consistency rc=$?" >> gpurun_out/summary_n$n.txt; tail -3 gpurun_out/dp_consistency31_n$n.log >> gpurun_out/summary_n$n.txt
cat gpurun_out/summary_n$n.txt
