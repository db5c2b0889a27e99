"""Static SASS summary of one kernel of an object file: splits the entry function at the
targets of its CALL instructions (the __noinline__ phase functions) and prints the opcode
mix of each part.  usage: sass_mix.py <obj> <substring of the mangled kernel name>"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
lines = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(lines) if "Function :" in l and pat in l)
end = start + 1
while end < len(lines) and "Function :" not in lines[end]:
    end += 1
ins = []
for l in lines[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
targets = sorted({int(re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", i).group(1), 16)
                  for a, i in ins if "CALL.REL.NOINC" in i})
cuts = [0] + targets + [ins[-1][0] + 16]


def cls(i):
    i = re.sub(r"^@!?U?P\d+\s+", "", i)
    return i.split()[0].split(".")[0]


for a, b in zip(cuts[:-1], cuts[1:]):
    seg = [i for ad, i in ins if a <= ad < b]
    c = collections.Counter(cls(i) for i in seg)
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print("%#7x n=%4d fp64=%4d lds=%3d sts=%3d shfl=%3d ldl/stl=%d/%d | %s" % (
        a, len(seg), fp, c["LDS"], c["STS"], c["SHFL"], c["LDL"], c["STL"],
        " ".join("%s:%d" % kv for kv in c.most_common(10))))
