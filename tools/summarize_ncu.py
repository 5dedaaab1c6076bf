"""Turn ncu outputs into the small, tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_tc_gemm_full.md [stage names...]
"""
import collections, csv, re, subprocess, sys


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum launch list ({src}): {sum(a[0] for a in agg.values())} launches, "
                f"{tot / 1e3:.2f} ms total (cold-cache, serialised: compare SHARES)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")


WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def full(src, dst, names):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"ncu --set full capture ({src}), per launch\n\n| stage | kernel | " + " | ".join(n for _, n in WANT) + " |\n|---|---|" + "---|" * len(WANT) + "\n")
        for i, r in enumerate(rows[2:]):
            vals = []
            for k, _ in WANT:
                if k not in hdr:
                    vals.append("n/a")
                    continue
                j = hdr.index(k)
                try:
                    x = float(r[j].replace(",", ""))
                    x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "ms": 1e3, "s": 1e6}.get(units[j], 1.0)
                    vals.append(f"{x:.1f}")
                except ValueError:
                    vals.append(r[j])
            nm = names[i] if i < len(names) else str(i)
            kname = r[ki].split("(")[0].replace("void ", "")[:48]
            f.write(f"| {nm} | `{kname}` | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
