# Round-end style validation on one B200: smoke, every GPU test, the bench line (all records), the reference arm, all models.
set -x
python __graft_entry__.py --smoke > gpurun_out/r03b_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r03b_tests.log
python bench.py > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03b_bench_reference.json 2>> gpurun_out/r03b_bench.err
python tools/bench_models.py > gpurun_out/models_r03b.jsonl 2> gpurun_out/r03b_models.err
tail -3 gpurun_out/r03b_smoke.log; tail -4 gpurun_out/r03b_tests.log; cut -c1-420 gpurun_out/r03b_bench.json; cut -c1-300 gpurun_out/r03b_bench_reference.json
python - <<'PY'
import json
for l in open("gpurun_out/models_r03b.jsonl"):
    d = json.loads(l)
    print(d["model"], round(d["frames_per_s"]), round(d["ms_per_batch"], 2))
PY
du -sh gpurun_out
