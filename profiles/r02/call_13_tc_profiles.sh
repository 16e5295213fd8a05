#!/bin/bash
mkdir -p gpurun_out/r02 /tmp/ncu
# the default bench line (never under a profiler)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_06.json 2> gpurun_out/r02/bench_06.err; echo bench rc=$?
tail -2 gpurun_out/r02/bench_06.err
# pipeline timeline of the tensor-core kernel (clock64 stamps)
timeout 100 python profiles/trace_chamfer_tc.py > gpurun_out/r02/trace_chamfer_tc.txt 2>&1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench_r02_tc.csv python bench.py --steps 2 --warmup 3 --no-configs --no-ref-gpu --no-cpu-baseline > /tmp/ncu/launches.out 2>&1
# full capture of the tensor-core Chamfer kernel (tf32 default and the fp16 variant)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chamfer_tc_kernel -s 2 -c 1 -o /tmp/ncu/tc2 -f python profiles/run_chamfer_tc_once.py 2 > /tmp/ncu/tc2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chamfer_tc_kernel -s 2 -c 1 -o /tmp/ncu/tc3 -f python profiles/run_chamfer_tc_once.py 3 > /tmp/ncu/tc3.log 2>&1
python profiles/ncu_summary.py /tmp/ncu/tc2.ncu-rep gpurun_out/r02/ncu_chamfer_tc_summary.csv > gpurun_out/r02/ncu_chamfer_tc_summary.txt
python profiles/ncu_summary.py /tmp/ncu/tc3.ncu-rep gpurun_out/r02/ncu_chamfer_tc_f16_summary.csv > gpurun_out/r02/ncu_chamfer_tc_f16_summary.txt
cp /tmp/ncu/tc2.ncu-rep gpurun_out/r02/ncu_chamfer_tc_final.ncu-rep
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/bench_06.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"])
r = d["roofline"]
print(r["ms_per_launch"], r["frac"], r["tensor_filter"], r["fp32_pipe_forms"]["one_cta_per_row_block"], r["fp32_pipe_forms"]["column_split_units"])
print(json.dumps(d.get("configs"))[:1500])
print(json.dumps(d.get("ref_gpu"))[:900])
PY
grep -E "Kernel Name|time_duration|tensor|alu_cycles|issue_active|dram__bytes|registers|top_stalls" gpurun_out/r02/ncu_chamfer_tc_summary.txt gpurun_out/r02/ncu_chamfer_tc_f16_summary.txt
