"""Instruction / stall-sample share per source region of a kernel in an ncu report.
   python tools/phase_share.py rep.ncu-rep kernel_regex"""
import collections, csv, io, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
# map (line no, text) -> file + enclosing function by scanning our sources
index = {}
for f in os.listdir(os.path.join(root, "rasterize_b200", "csrc")):
    fn = None
    for i, line in enumerate(open(os.path.join(root, "rasterize_b200", "csrc", f), errors="ignore"), 1):
        t = line.strip()
        if ("__device__" in t or "__global__" in t or t.startswith("template")) and not t.endswith(";"):
            pass
        import re
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__)?.*?\b([a-z_0-9]+)\s*\(", t) if ("__device__" in t or "__global__" in t or (t and t[0].isalpha() and t.endswith("{") and "(" in t and not t.startswith(("if", "for", "while", "else", "switch", "do", "const", "auto")))) else None
        if m:
            fn = m.group(1)
        index[(i, t[:60])] = (f, fn)
agg_i, agg_s = collections.Counter(), collections.Counter()
cur = None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0].isdigit():
        key = (int(r[0]), r[1].strip()[:60])
        cur = index.get(key, ("?", "?"))
        try:
            agg_i[cur] += int(r[7]); agg_s[cur] += int(r[6])
        except (ValueError, IndexError):
            pass
ti, ts = sum(agg_i.values()), sum(agg_s.values())
print(f"total warp-instr {ti}, samples {ts}")
for k, v in agg_i.most_common(14):
    print(f"{k[0]:20s} {str(k[1]):26s} inst {100 * v / ti:5.1f}%   samples {100 * agg_s[k] / max(ts, 1):5.1f}%")
