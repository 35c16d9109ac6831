cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
n=${NGPU:-8}
for c in 3 2; do
S3D_GRAD_CHUNKS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench33_n${n}_c$c.log 2>&1
echo "chunks=$c rc=$?"; grep "^{" gpurun_out/bench33_n${n}_c$c.log | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['n_gpus'], round(j['value']/1e6,1), 'M rays/s', round(j['ms_per_step'],3), 'ms', 'e2e', round(j['e2e']['ms_per_step'],3))"
done
