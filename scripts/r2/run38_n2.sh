cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "peer" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dp_peer_check.py > gpurun_out/r2/peer38_n2.log 2>&1
echo "rc=$?"; grep -v "Warning\|warn\|custom_\|^\*\|OMP_NUM" gpurun_out/r2/peer38_n2.log | tail -10
