"""profiles/traffic.json from an ncu --set full capture of k_apply_circ7 on the c4 mesh, stamped
with the hash of the kernel sources the capture was taken with (bench.py ignores stale entries).
usage: stamp_traffic.py report.ncu-rep details.txt   (run right after the capture, same tree)"""
import csv, io, json, os, subprocess, sys, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, details = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, r = rows[0], rows[1], rows[2]
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tot = 0
for nm in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = H.index(nm)
    tot += float(r[i]) * mult[U[i]]
with open(os.path.join(ROOT, details), "w") as fh:
    fh.write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
ent = {"c4": {"bytes": int(tot), "kernel": r[H.index("Kernel Name")][:40], "kernel_source_sha": b.kernel_source_sha(b.CIRC7_SOURCES),
              "source": "%s: dram__bytes_read.sum + dram__bytes_write.sum of one launch on the full c4 mesh, "
                        "ncu --set full --clock-control none" % details}}
json.dump(ent, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(ent))
