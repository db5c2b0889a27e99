"""Summarise the ptxas -v logs the Makefile leaves next to the objects
(radiative-transfer-amr_b200/csrc/*.o.ptxas.log) into profiles/r02_ptxas_summary.txt:
registers, stack and spill bytes of every entry kernel."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "radiative-transfer-amr_b200", "csrc")
rows = []
for log in sorted(glob.glob(os.path.join(CSRC, "*.o.ptxas.log"))):
    tu = os.path.basename(log).split(".")[0]
    text = open(log).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes "
                         r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", text):
        name, stack, st, ld, regs = m.groups()
        rows.append("%-40s regs %3s  stack %5s B  spill st/ld %5s/%5s B  %s"
                    % (tu, regs, stack, st, ld, name[:90]))
with open(os.path.join(ROOT, "profiles", "r02_ptxas_summary.txt"), "w") as fh:
    fh.write("ptxas -v summary (sm_100a), entry kernels only; device functions called through the ABI "
             "are listed in the build logs\n")
    fh.write("\n".join(rows) + "\n")
print(len(rows), "kernels")
