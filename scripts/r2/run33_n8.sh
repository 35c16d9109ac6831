cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/dp_peer_check.py > gpurun_out/r2/peer33_n8.log 2>&1
echo "peer check rc=$?"; grep -v "Warning\|warn\|custom_" gpurun_out/r2/peer33_n8.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench33_n8.log 2> gpurun_out/r2/bench33_n8.err
echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench33_n8.log').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches','clocks']}); print(d['e2e']); print(d['kernel_breakdown_ms_per_step'])
PY
