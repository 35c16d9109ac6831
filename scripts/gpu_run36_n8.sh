cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench36_n8.log 2>&1
echo "rc=$?"; grep "^{" gpurun_out/bench36_n8.log | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['n_gpus'], round(j['value']/1e6,1), 'M rays/s', round(j['ms_per_step'],3), 'ms', 'e2e', round(j['e2e']['ms_per_step'],3), round(j['e2e']['value']/1e6,1))"
