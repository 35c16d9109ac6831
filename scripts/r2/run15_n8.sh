cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2/bench15_n$N.log 2> gpurun_out/r2/bench15_n$N.err; echo "n$N rc=$?"
python scripts/bench_summary.py gpurun_out/r2/bench15_n$N.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/allreduce_bench.py > gpurun_out/r2/allreduce15_n8.log 2>&1; grep all_reduce gpurun_out/r2/allreduce15_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 scripts/dp_consistency.py > gpurun_out/r2/dp_consistency15_n8.log 2>&1; tail -3 gpurun_out/r2/dp_consistency15_n8.log
