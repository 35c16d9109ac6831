cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench10_n1.log 2> gpurun_out/r2/bench10_n1.err; echo "n1 rc=$?"
python scripts/bench_summary.py gpurun_out/r2/bench10_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench10_n2.log 2> gpurun_out/r2/bench10_n2.err; echo "n2 rc=$?"
python scripts/bench_summary.py gpurun_out/r2/bench10_n2.log
