"""Hot spots of an `ncu --page source --csv` SASS export: samples per execution-count bucket and the top instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
out = []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] in ('Kernel Name',):
        continue
    out.append(r)
ci, cs, cx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[cs]) for r in out)
totx = sum(int(r[cx]) for r in out)
print(f'{len(out)} SASS instructions, {totx} warp-instructions executed, {tot} samples')
b = collections.OrderedDict()
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for r in out:
    k = round(int(r[cx]) / nwarps)
    e = b.setdefault(k, [0, 0, 0])
    e[0] += 1; e[1] += int(r[cx]); e[2] += int(r[cs])
print('exec/warp  #sass  executed-share  sample-share')
for k, e in sorted(b.items(), key=lambda kv: -kv[1][2])[:12]:
    print(f'{k:9d} {e[0]:6d} {e[1] / totx:10.3f} {e[2] / max(tot, 1):10.3f}')
st = collections.Counter()
for r in out:
    for i, h in stall_cols:
        st[h] += int(r[i])
print('stalls:', ', '.join(f'{h[6:]}={v / max(tot, 1):.2f}' for h, v in st.most_common(8)))
print('top instructions by samples:')
for r in sorted(out, key=lambda r: -int(r[cs]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f'{int(r[cs]):6d} x{int(r[cx]) / nwarps:8.1f}  {r[ci].strip()[:90]}')
