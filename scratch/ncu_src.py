"""Summarise the ncu source page of one launch: top SASS instructions by stall samples.  usage: ncu_src.py rep launch_idx [top]"""
import csv, subprocess, sys, io
rep, li = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li), '--launch-count', '1'], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[start - 1][:100])
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
data = []
for n, r in enumerate(rows[1:]):
    try: s = int(r[ix['# Samples']])
    except: continue
    data.append((s, n, r))
tot = sum(d[0] for d in data)
print('total samples', tot)
for s, n, r in sorted(data, reverse=True)[:top]:
    st = sorted([(int(r[ix[k]] or 0), k) for k in stalls], reverse=True)[:2]
    print('%6d %5.1f%%  #%4d  %-70s %s' % (s, 100.0 * s / tot, n, r[ix['Source']].strip()[:70], ' '.join('%s=%d' % (k[6:], v) for v, k in st)))
