"""Top SASS instructions by stall samples from an ncu report: python tools/sass_hot.py rep.ncu-rep kernel_regex [n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [r for r in rows if r and r[0] == "Line No"][0]
ix = {h: i for i, h in enumerate(hdr)}
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sass, cur, seen = [], None, set()
for r in rows:
    if not r:
        continue
    if r[0].isdigit():
        cur = (int(r[0]), r[1].strip()[:70])
        continue
    if len(r) > 8 and r[2].startswith("0x") and r[2] not in seen:
        seen.add(r[2])
        try:
            sass.append((int(r[6]), int(r[7]), r[3].strip(), cur, r))
        except ValueError:
            pass
tot = sum(x[0] for x in sass)
texec = sum(x[1] for x in sass)
print("samples", tot, "warp-instr", texec)
agg = {}
for x in sass:
    for h in stallcols:
        agg[h] = agg.get(h, 0) + int(x[4][ix[h]] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items()), reverse=True)[:6])
for x in sorted(sass, key=lambda x: -x[0])[:n]:
    r = x[4]
    stalls = {h: int(r[ix[h]] or 0) for h in stallcols}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * x[0] / max(tot, 1):5.2f}% exec={x[1]:7d} {x[2][:52]:52s} L{x[3][0]:<4d} {top[0][0]}={top[0][1]} | {x[3][1][:50]}")
