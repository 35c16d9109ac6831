cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
# exactly the driver's launch (reference arm first, then ours), N = 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench03_ref_n2.log 2> gpurun_out/r2/bench03_ref_n2.err; echo "ref rc=$?"
tail -c 400 gpurun_out/r2/bench03_ref_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench03_n2.log 2> gpurun_out/r2/bench03_n2.err; echo "ours rc=$?"
tail -c 800 gpurun_out/r2/bench03_n2.err
python scripts/bench_summary.py gpurun_out/r2/bench03_n2.log
