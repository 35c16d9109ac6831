cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
L=gpurun_out/r2/stepbench04.log; : > $L
python scripts/r2/stepbench.py --tag "baseline chunks=1" --breakdown >> $L 2>&1
for C in 2 4 8; do for B in 256 128; do
S3D_BWD_CHUNKS=$C S3D_SCATTER_BLOCK=$B python scripts/r2/stepbench.py --tag "bwd_chunks=$C scatter_block=$B" >> $L 2>&1
done; done
S3D_BWD_CHUNKS=4 python scripts/r2/stepbench.py --tag "bwd_chunks=4 no-prefetch" --no-prefetch >> $L 2>&1
python scripts/r2/stepbench.py --tag "chunks=1 no-prefetch" --no-prefetch >> $L 2>&1
cat $L
