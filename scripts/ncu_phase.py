"""Per-phase summary of an ncu source-level capture (SASS view): instructions per warp, stall
samples and shared-memory wavefronts between the kernel's block barriers, plus the opcode mix.
usage: ncu_phase.py report.ncu-rep [launch index] [blocks]"""
import csv, re, collections, subprocess, sys, io
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:80])
H = rows[1]
ci = H.index('Instructions Executed'); cs = H.index('# Samples'); cw = H.index('L1 Wavefronts Shared')
stall = {h: i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h}
data = []
for r in rows[2:]:
    try: data.append((r[1], float(r[ci] or 0), float(r[cs] or 0), float(r[cw] or 0), r))
    except Exception: pass
# the export lists the function twice when it has two entry variants: keep the first copy
half = len(data)
names = [d[0] for d in data]
for k in range(len(data) // 2 - 2, len(data) // 2 + 3):
    if k > 0 and names[:k] == names[k:2 * k]:
        half = k
data = data[:half] if half != len(data) else data
tot_i = sum(d[1] for d in data); tot_s = sum(d[2] for d in data)
print("sass lines %d, warp-instructions %.3e, samples %d" % (len(data), tot_i, tot_s))
op = collections.defaultdict(lambda: [0, 0])
for d in data:
    t = re.sub(r'^\s*(@!?U?P\d+\s+)?', '', d[0]).split()
    o = t[0].split('.')[0] if t else '?'
    op[o][0] += d[1]; op[o][1] += d[2]
print("opcode mix:", ", ".join("%s %.1f%%/%.1f%%" % (o, 100 * v[0] / tot_i, 100 * v[1] / tot_s)
                                for o, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:14]))
cuts = [0] + [k for k, d in enumerate(data) if 'BAR.SYNC' in d[0]] + [len(data)]
for a, b in zip(cuts[:-1], cuts[1:]):
    seg = data[a:b]
    i = sum(d[1] for d in seg); s = sum(d[2] for d in seg); w = sum(d[3] for d in seg)
    st = collections.Counter()
    for d in seg:
        for h, c in stall.items():
            try: st[h[6:]] += float(d[4][c] or 0)
            except Exception: pass
    dm = sum(d[1] for d in seg if 'DMMA' in d[0])
    print("[%5d,%5d) inst %5.1f%% samples %5.1f%% smem-wavefronts %.3e dmma %.3e | %s" % (
        a, b, 100 * i / tot_i, 100 * s / max(tot_s, 1), w, dm, ", ".join("%s %d" % kv for kv in st.most_common(5))))
