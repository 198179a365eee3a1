set -x
mkdir -p /tmp/o
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "lstm_cell" 2>&1 | tail -4 > gpurun_out/r02m_tests.log
python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "fullsubnet or dpcrn or plan_abi" 2>&1 | tail -4 >> gpurun_out/r02m_tests.log
python tools/bench_models.py fullsubnet > gpurun_out/models_r02m_fsn_ew16.jsonl 2>gpurun_out/r02m.err
SE_CELL_EPI_WARPS=8 python tools/bench_models.py fullsubnet > gpurun_out/models_r02m_fsn_ew8.jsonl 2>>gpurun_out/r02m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02m.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /tmp/o/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_pair -s 1 -c 1 -o /tmp/o/ncu_gemm_f16 python tools/gemm_once.py 3 f16 > /tmp/o/ncu_gemm.log 2>&1
ls -la /tmp/o gpurun_out
python tools/ncu_summarize.py full /tmp/o/ncu_gemm_f16.ncu-rep gpurun_out/ncu_gemm_f16_r02m.json "ncu --set full --clock-control none -k regex:gemm_tf32x3_pair -s 1 -c 1 python tools/gemm_once.py 3 f16" > /dev/null 2>>gpurun_out/r02m.err
ncu --set full --clock-control none -k regex:conv_tf32x3_kernel -s 16 -c 16 -o /tmp/o/ncu_conv_f16 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /tmp/o/ncu_conv.log 2>&1
ncu -i /tmp/o/ncu_conv_f16.ncu-rep --page raw --csv > /tmp/o/conv_raw.csv 2>>gpurun_out/r02m.err
python - <<'PY' > gpurun_out/ncu_conv_f16_r02m_summary.txt 2>>gpurun_out/r02m.err
import csv
rows=list(csv.reader(open('/tmp/o/conv_raw.csv')))
hdr=rows[0]
keep=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","launch__grid_size","launch__registers_per_thread"]
idx=[hdr.index(k) for k in keep if k in hdr]
print("ncu --set full --clock-control none -k regex:conv_tf32x3_kernel -s 16 -c 16 python bench.py --steps 1 --warmup 1  (one CRN step, fp16-pair convs)")
print(" | ".join(hdr[i] for i in idx))
print(" | ".join(rows[1][i] for i in idx))
for r in rows[2:]:
    print(" | ".join(r[i][:60] for i in idx))
PY
tail -3 gpurun_out/r02m_tests.log; cut -c1-330 gpurun_out/models_r02m_fsn_ew16.jsonl gpurun_out/models_r02m_fsn_ew8.jsonl; du -sh gpurun_out
