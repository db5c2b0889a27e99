"""Per-function summary of an ncu source-page CSV export (ncu -i rep --page source --csv
--print-source sass > file.csv): the kernel is cut at the targets of its CALL instructions
(the __noinline__ phase functions); prints instructions, stall samples, shared-memory
wavefronts per part and the barrier lines.  usage: ncu_funcs.py file.csv [units]
(units: what to divide the counts by, e.g. the number of columns)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
H = rows[1]
ci, cs, cw = H.index('Instructions Executed'), H.index('# Samples'), H.index('L1 Wavefronts Shared')
stall = {h: i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h}
data = rows[2:]
addr = [int(r[0], 16) for r in data]
base = addr[0]
calls = sorted({int(re.search(r'CALL\.REL\.NOINC\s+(0x[0-9a-f]+)', r[1]).group(1), 16)
                for r in data if 'CALL.REL.NOINC' in r[1]})
cuts = [base] + calls + [addr[-1] + 16]
f = lambda v: float(v or 0)
tot_s = sum(f(r[cs]) for r in data)
tot_i = sum(f(r[ci]) for r in data)
print('samples %d, warp-instructions %.4g (%.0f per unit)' % (tot_s, tot_i, tot_i / units))
for a, b in zip(cuts[:-1], cuts[1:]):
    seg = [r for r, ad in zip(data, addr) if a <= ad < b]
    i = sum(f(r[ci]) for r in seg)
    if i == 0:
        continue
    s = sum(f(r[cs]) for r in seg)
    w = sum(f(r[cw]) for r in seg)
    st = collections.Counter()
    for r in seg:
        for h, c in stall.items():
            st[h[6:]] += f(r[c])
    print("%#8x n=%4d inst/unit %6.0f (%4.1f%%) samples %5.1f%% smem wf/unit %6.0f | %s" % (
        a - base, len(seg), i / units, 100 * i / tot_i, 100 * s / tot_s, w / units,
        ", ".join("%s %d" % kv for kv in st.most_common(6))))
print("barrier waits (samples at the instruction after each BAR):")
for k, r in enumerate(data[:-1]):
    if 'BAR.SYNC' in r[1] or 'BAR.ARV' in r[1]:
        nxt = data[k + 1]
        print("  line %4d %-40s executed %10.0f  -> %6.0f samples (%.1f%%)" % (
            k, r[1].strip()[:40], f(r[ci]), f(nxt[cs]), 100 * f(nxt[cs]) / tot_s))
