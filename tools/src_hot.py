"""Per-file / per-line warp-instruction and stall-sample shares of one kernel of an ncu report.

    python tools/src_hot.py gpurun_out/prof.ncu-rep [kernel-regex] [top]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
by_samples = len(sys.argv) > 4 and sys.argv[4] == "samples"  # order the lines by stall samples instead of instructions
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre:
    cmd += ["--kernel-name", "regex:" + kre]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
cur = None
per_file = collections.Counter()
per_file_s = collections.Counter()
lines = collections.defaultdict(lambda: [0, 0, ""])
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] in ("", "-"):  # a source-line row (its SASS rows follow, with an address)
        try:
            n, s = int(r[7]), int(r[6])
        except ValueError:
            continue
        k = (cur, int(r[0]))
        lines[k][0] += n
        lines[k][1] += s
        lines[k][2] = r[1].strip()[:100]
        per_file[cur] += n
        per_file_s[cur] += s
tot = sum(per_file.values()) or 1
tots = sum(per_file_s.values()) or 1
print(f"# {rep}: {tot} warp-instructions, {tots} stall samples")
for f, n in per_file.most_common():
    print(f"{100 * n / tot:5.1f}% instr  {100 * per_file_s[f] / tots:5.1f}% samples  {f}")
print()
for (f, ln), (n, s, text) in sorted(lines.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print(f"{100 * n / tot:5.1f}% instr {100 * s / tots:5.1f}% smp  {f}:{ln:<4d} {text}")
