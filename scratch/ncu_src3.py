import csv, subprocess, sys, io, collections
rep, li, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li), '--launch-count', '1'], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = collections.Counter(); n_inst = 0; ex = 0; samples = 0
for n, r in enumerate(rows[1:]):
    if n < lo or n > hi: continue
    try: s = int(r[ix['# Samples']])
    except: continue
    samples += s
    e = int(r[ix['Instructions Executed']] or 0); ex += e
    if e: n_inst += 1
    for k in stalls: tot[k] += int(r[ix[k]] or 0)
print('instructions with executions:', n_inst, 'warp-instr executed:', ex, 'samples', samples)
print(tot.most_common(8))
