# Source-level stall samples of the fp16-pair conv kernel inside one CRN step (development evidence; run under gpurun).
set -x
mkdir -p /tmp/o
ncu --set full --clock-control none --import-source on -k regex:conv_tf32x3_kernel -s 9 -c 9 -o /tmp/o/conv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /tmp/o/log 2>&1
ls -la /tmp/o
for i in 1 8; do
  ncu -i /tmp/o/conv.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > /tmp/o/src_$i.csv 2>/tmp/o/err_$i
  python - "$i" <<'PY'
import csv, sys
i = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/o/src_{i}.csv")))
hdr = None
out = []
for r in rows:
    if "Source" in r and hdr is None:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        out.append(r)
print(i, "header", hdr[:40] if hdr else None, "rows", len(out))
if hdr:
    def col(name):
        return hdr.index(name) if name in hdr else None
    cs, csamp, cinst = col("Source"), col("# Samples") or col("Samples"), col("Instructions Executed")
    keep = [c for c in (col("Address"), cs, csamp, cinst, col("Warp Stall Sampling (All Samples)"), col("stall_long_sb"), col("stall_wait"), col("stall_barrier"), col("stall_short_sb"), col("stall_math")) if c is not None]
    with open(f"gpurun_out/conv_src_{i}.csv", "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in keep])
        for r in out:
            w.writerow([r[c][:90] for c in keep])
PY
  head -c 600 /tmp/o/src_$i.csv
done
ncu -i /tmp/o/conv.ncu-rep --page raw --csv > /tmp/o/raw.csv
python - <<'PY'
import csv
rows = list(csv.reader(open("/tmp/o/raw.csv")))
hdr = rows[0]
want = [h for h in hdr if any(k in h for k in ("Kernel Name", "gpu__time_duration.sum", "smsp__pcsamp_warps_issue_stalled", "sm__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct", "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")) and "not_issued" not in h]
idx = [hdr.index(h) for h in want]
with open("gpurun_out/conv_raw_r02p.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(want)
    for r in rows[1:]:
        w.writerow([r[i][:70] for i in idx])
PY
du -sh gpurun_out
