cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for b in 64 96 128 192; do S3D_SCATTER_BLOCK=$b python scripts/r2/stepbench.py --tag "scatter block $b" --breakdown 2>&1 | tail -1 | cut -c1-200; done
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "scatter or backward or trainer or deterministic" 2>&1 | tail -2
