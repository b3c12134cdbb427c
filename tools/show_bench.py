"""print the interesting parts of a bench.py JSON line (development helper)"""
import json, sys
for line in open(sys.argv[1]):
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print('main %.1f %s  %.3f ms/step  e2e %.1f (%.2f ms, d2h %.0f MB)  roof %.3f  launches %d' % (
        d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'] / 1e6,
        d['roofline']['frac'], d['gpu_launches']))
    print('  stages', d.get('stages'))
    print('  cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d.get('clocks'))
    for k, v in d.get('secondary', {}).items():
        if 'error' in v:
            print(' ', k, v)
            continue
        print('  %-10s %10.1f %-8s %8.3f ms  e2e %10.1f  roof %s %s  cpu %s' % (
            k, v['value'], v['unit'], v['ms_per_step'], v['e2e']['value'],
            v['roofline'] and round(v['roofline']['frac'], 3), [round(r['frac'], 3) for r in v.get('roofline_more', [])],
            v.get('cpu_baseline', {}).get('value')))
        if v.get('stats'):
            print('             ', v['stats'])
