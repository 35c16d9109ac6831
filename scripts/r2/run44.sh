cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/r2/bench44.log 2> gpurun_out/r2/bench44.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2/bench44.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'])"
