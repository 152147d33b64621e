"""Groups the per-instruction page of an .ncu-rep (`ncu --set full --import-source on`) by execution count, which
separates the phases of a persistent kernel (inner loop / per-source set-up / per-batch prologue / ...), and prints
for each phase: static instructions, share of executed instructions, share of warp-stall samples, shared-memory
wavefronts, and the dominant stall reasons. Optionally lists the instructions of one phase.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep [exec_count_to_list]
"""
import csv
import subprocess
import sys
from collections import defaultdict


def load(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    return rows[hi], rows[hi + 1:]


def main(path, list_count=None):
    hdr, data = load(path)
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    wf_col = ix.get("L1 Wavefronts Shared")

    def num(r, col):
        try:
            return int(r[col] or 0)
        except (ValueError, IndexError):
            return 0

    tot_s = sum(num(r, ix["# Samples"]) for r in data) or 1
    tot_e = sum(num(r, ix["Instructions Executed"]) for r in data) or 1
    groups = defaultdict(lambda: {"n": 0, "samples": 0, "exec": 0, "wf": 0, "stall": defaultdict(int)})
    for r in data:
        g = groups[num(r, ix["Instructions Executed"])]
        g["n"] += 1
        g["samples"] += num(r, ix["# Samples"])
        g["exec"] += num(r, ix["Instructions Executed"])
        g["wf"] += num(r, wf_col) if wf_col is not None else 0
        for h in stalls:
            g["stall"][h[6:]] += num(r, ix[h])
    print(f"{len(data)} instructions, {tot_e} executed (warp level), {tot_s} stall samples")
    print("exec/instr  static  inst%  samples%  smem wavefronts  top stalls")
    for count, g in sorted(groups.items(), key=lambda kv: -kv[1]["exec"])[:16]:
        top = sorted(g["stall"].items(), key=lambda kv: -kv[1])[:4]
        print(f"{count:10d} {g['n']:7d} {100 * g['exec'] / tot_e:6.2f} {100 * g['samples'] / tot_s:9.2f} {g['wf']:16d}  "
              + " ".join(f"{k}={v}" for k, v in top if v))
    if list_count is not None:
        for k, r in enumerate(data):
            if num(r, ix["Instructions Executed"]) == list_count:
                st = sorted(((h[6:], num(r, ix[h])) for h in stalls), key=lambda kv: -kv[1])[:2]
                print(k, r[ix["Source"]].strip()[:90].ljust(90), r[ix["# Samples"]].rjust(5), " ".join(f"{a}={b}" for a, b in st if b))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
