"""Turn Nsight Compute output brought back in gpurun_out/ into the small, committed files under profiles/.

    python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/launches_rNN_summary.txt "<command line>"
    python tools/ncu_summarize.py full gpurun_out/ncu_full.ncu-rep profiles/ncu_full_rNN.json "<command line>"

`launches`: per-kernel totals / shares of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised:
compare shares, not absolutes).  `full`: the handful of metrics DESIGN.md and bench.py quote from a `--set full`
capture (runs `ncu -i <rep> --page raw --csv` here; no GPU needed).
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("se::", "").strip()


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        if v != v:
            continue
        ms = v / 1e6 if r[iu] == "ns" else (v / 1e3 if r[iu] in ("us", "usecond") else v)
        k = short(r[ik])[:110]
        c, t = tot.get(k, (0, 0.0))
        tot[k] = (c + 1, t + ms)
    total = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write(f"{cmd}\n(cold-cache, serialised launches: compare SHARES; total {total:.3f} ms over "
                f"{sum(c for c, _ in tot.values())} launches)\n")
        for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.3f} ms {c:5d} launches {100 * t / total:5.1f}%  {k}\n")
    print(open(dst).read())


def full(src, dst, cmd):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": short(r[hdr.index("Kernel Name")])}
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                try:
                    rec[m] = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                rec[m + "__unit"] = units[i]
        if rec.get("gpu__time_duration.sum", float("nan")) == rec.get("gpu__time_duration.sum"):
            out.append(rec)
    json.dump({"source": cmd, "kernels": out}, open(dst, "w"), indent=1)
    # profiles/roofline_traffic.json: DRAM bytes per launch of every kernel captured so far, newest capture wins --
    # the ONE place bench.py reads `roofline.traffic` from (so the number can never come from a stale kernel's capture)
    import os
    tpath = os.path.join(os.path.dirname(os.path.abspath(dst)), "roofline_traffic.json")
    table = json.load(open(tpath)) if os.path.exists(tpath) else {}
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for k in out:
        if "dram__bytes_read.sum" in k and "dram__bytes_write.sum" in k:
            tot = (k["dram__bytes_read.sum"] * unit_scale.get(k.get("dram__bytes_read.sum__unit", "byte"), 1.0) +
                   k["dram__bytes_write.sum"] * unit_scale.get(k.get("dram__bytes_write.sum__unit", "byte"), 1.0))
            name = re.sub(r"<.*", "", k["kernel"]).strip()
            table[name] = {"dram_bytes_per_launch": tot, "source": os.path.basename(dst)}
    json.dump(table, open(tpath, "w"), indent=1, sort_keys=True)
    for k in out:
        print(f"{k['kernel'][:40]:40s} {k.get('gpu__time_duration.sum', 0):8.3f} {k.get('gpu__time_duration.sum__unit', 'ms')}  DRAM r+w "
              f"{k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0):8.1f} "
              f"{k.get('dram__bytes_read.sum__unit', '')}  dram% "
              f"{k.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}  tensor% "
              f"{k.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}  regs "
              f"{k.get('launch__registers_per_thread', 0):.0f}  L2hit% {k.get('lts__t_sector_hit_rate.pct', 0):.0f}")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    (launches if mode == "launches" else full)(src, dst, cmd)
