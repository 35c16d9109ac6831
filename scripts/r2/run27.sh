cd $GRAFT_REPO_ROOT
timeout 300 python scripts/r2/wide_trace.py 2>&1 | tail -40
