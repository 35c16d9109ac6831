cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench16_n8.log 2> gpurun_out/r2/bench16_n8.err; echo "n8 rc=$?"
python scripts/bench_summary.py gpurun_out/r2/bench16_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench16_ref_n8.log 2> gpurun_out/r2/bench16_ref_n8.err; echo "ref n8 rc=$?"; tail -c 300 gpurun_out/r2/bench16_ref_n8.log
