cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dp_peer_check.py > gpurun_out/r2/peer32_n2.log 2>&1
echo "rc=$?"; grep -v "Warning\|warn" gpurun_out/r2/peer32_n2.log | tail -30
