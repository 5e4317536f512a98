"""Summarise an `ncu --page source --csv` dump: stall reasons, executed opcode mix, hottest instructions.
    python tools/ncu_sass_report.py <sass.csv> [rows_per_launch]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
i_src, i_ex, i_samp, i_addr = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
st, mix, tot, data = collections.Counter(), collections.Counter(), 0, []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    op = [o for o in r[i_src].split() if not o.startswith('@')][0].split('.')[0]
    n = int(r[i_ex]); tot += n; mix[op] += n
    s = {hdr[c]: int(r[c]) for c in stall_cols}
    for k, v in s.items():
        st[k] += v
    data.append((r[i_addr], r[i_src], n, int(r[i_samp]), s))
print('total warp-instructions', tot, ' per row-step %.1f' % (tot / per))
print('stalls', st.most_common(12))
for op, n in mix.most_common(32):
    print(f'{op:10s} {n:11d} {100*n/tot:5.1f}%  per-row {n/per:6.1f}')
print('--- top sampled instructions (samples, executed, sass, top stalls)')
for d in sorted(data, key=lambda d: -d[3])[:45]:
    top = sorted(d[4].items(), key=lambda kv: -kv[1])[:2]
    print(d[3], d[2], d[1][:72], top)
