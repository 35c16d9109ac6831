cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 scripts/dp_peer_check.py > gpurun_out/r2/peer39_n4.log 2>&1
echo "peer check rc=$?"; grep -v "Warning\|warn\|custom_\|^\*\|OMP_NUM" gpurun_out/r2/peer39_n4.log | tail -9
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench39_n4.log 2> gpurun_out/r2/bench39_n4.err
echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench39_n4.log').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches','clocks']}); print(d['e2e']); print(d['kernel_breakdown_ms_per_step'])
PY
