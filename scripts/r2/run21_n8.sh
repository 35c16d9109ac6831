cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for C in 1 2 3; do
S3D_GRAD_CHUNKS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$C bench.py --gpus 8 --steps 20 --warmup 5 --no-roofline > gpurun_out/r2/bench21_n8_c$C.log 2> gpurun_out/r2/bench21_n8_c$C.err; echo "chunks=$C rc=$?"
python scripts/bench_summary.py gpurun_out/r2/bench21_n8_c$C.log
done
