# ncu evidence on the final tree of round 2 (run under gpurun; summaries are made on the box, the reports stay in /tmp)
set -x
mkdir -p /tmp/o
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /tmp/o/ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tf32x3_pair -s 1 -c 1 -o /tmp/o/ncu_gemm python tools/gemm_once.py 3 f16 > /tmp/o/ncu_gemm.log 2>&1
python tools/ncu_summarize.py full /tmp/o/ncu_gemm.ncu-rep gpurun_out/ncu_gemm_f16_r02_final.json "ncu --set full --clock-control none -k regex:gemm_tf32x3_pair -s 1 -c 1 python tools/gemm_once.py 3 f16" > /dev/null 2>gpurun_out/ev.err
ncu --set full --clock-control none -k regex:conv_tf32x3_kernel -s 9 -c 9 -o /tmp/o/ncu_conv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /tmp/o/ncu_conv.log 2>&1
ncu -i /tmp/o/ncu_conv.ncu-rep --page raw --csv > /tmp/o/conv_raw.csv 2>>gpurun_out/ev.err
python - <<'PY' > gpurun_out/ncu_conv_f16_r02_final_summary.txt 2>>gpurun_out/ev.err
import csv
rows = list(csv.reader(open('/tmp/o/conv_raw.csv')))
hdr = rows[0]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
idx = [hdr.index(k) for k in keep if k in hdr]
print("ncu --set full --clock-control none -k regex:conv_tf32x3_kernel -s 9 -c 9 python bench.py --steps 1 --warmup 1   (the nine fp16-pair conv launches of one CRN step, final tree: fast ELU, batched bias loads, 256-bit stores)")
print(" | ".join(hdr[i] for i in idx))
print(" | ".join(rows[1][i] for i in idx))
for r in rows[2:]:
    print(" | ".join(r[i][:56] for i in idx))
PY
python tools/gemm_once.py 20 f16 > gpurun_out/gemm_once_r02_final.txt 2>&1
python tools/gemm_once.py 20 tf32 >> gpurun_out/gemm_once_r02_final.txt 2>&1
cat gpurun_out/gemm_once_r02_final.txt; du -sh gpurun_out
