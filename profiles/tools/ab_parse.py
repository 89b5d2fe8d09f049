import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith('{'):
        print(l[:200]); continue
    d = json.loads(l)
    r = d['roofline']
    print("fps %.1f ms %.1f  fullres %.0f GB/s frac %.3f" % (d['value'], d['ms_per_step'], r['achieved'], r['frac']))
    for k, v in r['per_class'].items():
        print("   %-16s %8.2f ms %6.0f GB/s" % (k, v['ms_per_step'], v['GBps']))
