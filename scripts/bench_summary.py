import json, sys
for f in sys.argv[1:]:
    for ln in open(f):
        if ln.startswith('{'):
            j = json.loads(ln)
            print(f, 'rays %d value %.3e ms %.3f e2e %.3e e2e_ms %.3f' % (j['config'].get('rays_per_step_per_gpu', 0), j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e'].get('ms_per_step', 0)))
            kb = j.get('kernel_breakdown_ms_per_step')
            if kb: print('  ', {k.replace('s3d_', ''): v for k, v in kb.items() if v > 0.02})
            r = j.get('roofline')
            if r: print('   roofline', r['kernel'], 'frac %.3f' % (r.get('frac') or -1), 'share %.2f' % (r.get('share_of_step') or -1), 'launch_ms %.3f' % r.get('launch_ms', 0))
            g = j.get('grid_encode_forward_roofline')
            if g: print('   grid_encode_forward: fp32 frac %.3f (%.0f GB/s) fp16 frac %.3f' % (g['fp32']['frac'], g['fp32']['achieved'], g['fp16']['frac']))
            if 'cpu_baseline' in j: print('   cpu_baseline %.1f rays/s on %d cores' % (j['cpu_baseline']['value'], j['cpu_baseline']['cores']))
