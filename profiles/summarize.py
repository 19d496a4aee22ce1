"""Summarise ncu artefacts brought back in gpurun_out/ into small text files that can be committed.

  python profiles/summarize.py launches gpurun_out/launches_c2.csv  > profiles/rNN_launches.txt
  python profiles/summarize.py full     gpurun_out/prof.ncu-rep      > profiles/rNN_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_atom.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        agg[re.sub(r"\(.*", "", row["Kernel Name"])].append(float(row["Metric Value"]))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {path}  (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:72]:72s} n={len(v):4d} mean={sum(v) / len(v) / 1000:9.2f} us  share={100 * sum(v) / total:5.1f}%")


def full(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h]
    print(f"# ncu --set full summary of {path}")
    seen = set()
    for d in data:
        name = re.sub(r"\(.*", "", d[idx["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        print(f"\n## {name}")
        for k in KEYS:
            if k in idx:
                print(f"{k:66s} {d[idx[k]]:>18s} {units[idx[k]]}")
        st = sorted(((float(d[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:6]
        print("top stalls (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in st))
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + re.escape(name.split("::")[-1].split("<")[0])],
                             capture_output=True, text=True).stdout
        per, tot = [], 0
        for r in csv.reader(io.StringIO(src)):
            if len(r) >= 8 and r[0].isdigit():
                try:
                    n, s = int(r[7]), int(r[6])
                except ValueError:
                    continue
                per.append((n, s, int(r[0]), r[1].strip()[:100]))
                tot += n
        # the capture may hold several launches of the kernel: report shares
        print(f"hottest source lines (share of {tot} warp-instructions in the capture):")
        for n, s, ln, text in sorted(per, reverse=True)[:top]:
            print(f"  {100 * n / max(tot, 1):5.1f}%  samples={s:6d}  L{ln:<4d} {text}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
