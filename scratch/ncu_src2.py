import csv, subprocess, sys, io
rep, li = sys.argv[1], int(sys.argv[2]); thr = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li), '--launch-count', '1'], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
acc = 0
for n, r in enumerate(rows[1:]):
    try: s = int(r[ix['# Samples']])
    except: continue
    src = r[ix['Source']].strip()
    key = any(t in src for t in ('SYNCS', 'UTCHMMA', 'UTCBAR', 'LDTM', 'BAR.', 'UBLKCP', 'EXIT', 'LDG', 'STG', 'STS', 'ELECT', 'UTCATOM'))
    acc += s
    if s >= thr or (key and s >= 10):
        st = sorted([(int(r[ix[k]] or 0), k) for k in stalls], reverse=True)[:2]
        print('#%5d cum %6d  %6d  %-64s %s' % (n, acc, s, src[:64], ' '.join('%s=%d' % (k[6:], v) for v, k in st)))
