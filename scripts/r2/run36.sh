cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 120 scripts/r2/bin/red_micro | tee gpurun_out/r2/red_micro36.log
