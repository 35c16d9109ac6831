cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python scripts/r2/zero_frac.py 2>&1 | grep samples | awk 'NR%4==1' 
SHAPE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/wide_ncu25.csv python scripts/r2/wide_micro.py > gpurun_out/r2/wide_micro25.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2/wide_ncu25.csv')) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: agg[d['Kernel Name'][:60]+' grid='+d.get('Grid Size','')].append(float(d['Metric Value'].replace(',','')))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print('%8.1f us x%d  %s'%(sum(v)/len(v), len(v), k))
PY
