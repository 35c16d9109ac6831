cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "deterministic or map_color or color or grid_encode_backward or scatter" > gpurun_out/r2/pt23.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt23.log | tail -30
python scripts/r2/stepbench.py --tag "default" --breakdown > gpurun_out/r2/stepbench23.log 2>&1; cat gpurun_out/r2/stepbench23.log
S3D_DETERMINISTIC=1 python scripts/r2/stepbench.py --tag "deterministic (fixed-point)" --breakdown > gpurun_out/r2/stepbench23_det.log 2>&1; cat gpurun_out/r2/stepbench23_det.log
timeout 900 python scripts/ref_ab.py --out gpurun_out/r2/ref_vs_ours23.json > gpurun_out/r2/ref_ab23.log 2>&1; tail -25 gpurun_out/r2/ref_ab23.log
