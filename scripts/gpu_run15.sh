cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/trace_bwd.py > gpurun_out/trace_bwd.log 2>&1; echo "trace rc=$?"
tail -60 gpurun_out/trace_bwd.log
