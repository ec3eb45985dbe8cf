"""Turn ncu outputs into the small CSV / JSON summaries kept under profiles/ (development aid).
  python tools/ncu_summarise.py launches <launch-list.csv> <out.csv> "<command>"
  python tools/ncu_summarise.py full <report.ncu-rep> <out.csv>          (needs ncu on PATH)"""
import collections, csv, io, json, re, subprocess, sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor", "sm__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_fp64", "sm__pipe_fp64",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "sm__inst_executed.sum", "smsp__inst_executed.sum", "local_load", "local_store",
        "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate"]


def short(name):
    name = re.sub(r"^void (unnamed>::|mpsb::[^:]*::)?", "", name)
    return name.split("(")[0]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        k = short(r[ik])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[iv].replace(",", "")) / 1e3
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"kernel,launches,total_us,share  ({cmd})\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / total:.4f}\n")


def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for i, h in enumerate(hdr):
            if any(k in h for k in KEEP):
                w.writerow([h, units[i]] + [d[i] for d in data])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
