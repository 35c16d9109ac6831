cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/r2/peer_diag.py 2>&1 | grep -v "Warning\|warn\|custom_\|^\*\|OMP_NUM" | tail -5
