"""Join an ncu SASS source page (csv) with nvdisasm -g line info: stall samples per CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> <cubin-prefix e.g. k4_ppo_lag> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, cub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "icrl_b200", "libicrl_b200.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
lines_of, cur, infn, n = [], None, False, 0
for l in dis.split("\n"):
    if l.startswith(".text.") and l.endswith(":"):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines_of.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be in the report; take the first whose name matches
out, i = {}, 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        if "ppo" in rows[i][1] or kern[:8] in rows[i][1] or True:
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            base = int(body[0][0], 16)
            for r in body:
                off = (int(r[0], 16) - base) // 16
                key = lines_of[off] if off < len(lines_of) else None
                s, e = out.get(key, (0, 0))
                out[key] = (s + int(r[si] or 0), e + int(r[ii] or 0))
            break
        i = j
    else:
        i += 1
tot = sum(s for s, _ in out.values())
tote = sum(e for _, e in out.values())
print(f"total samples {tot}, warp instructions {tote}")
srcs = {}
for (k, (s, e)) in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        f = os.path.join(root, "icrl_b200", "csrc", k[0])
        if os.path.exists(f):
            srcs.setdefault(f, open(f).read().split("\n"))
            text = srcs[f][k[1] - 1].strip()[:100]
    print(f"{100*s/tot:5.1f}% samples {100*e/max(tote,1):5.1f}% inst  {k}  {text}")
