cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench37_n2.log 2> gpurun_out/r2/bench37_n2.err
echo "bench n2 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2/bench37_ref_n2.log 2> gpurun_out/r2/bench37_ref_n2.err
echo "reference arm n2 rc=$?"; tail -c 600 gpurun_out/r2/bench37_ref_n2.log
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench37_n2.log').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches','clocks']}); print(d['e2e']); print(d['config']['parallelism']); print(d['kernel_breakdown_ms_per_step'])
PY
