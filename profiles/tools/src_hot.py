import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot)
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i in sorted(top):
    r = data[i]
    n = int(r[ix['# Samples']] or 0)
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("%5d %5.1f%% [%4d] %-70s %s" % (n, 100.0 * n / tot, i, r[ix['Source']].strip()[:70], " ".join("%s=%d" % (s[6:], v) for v, s in st if v)))
