mkdir -p gpurun_out/r03
for s in 1 0 1 0 2; do
  PDAE_CHAMFER_SPLIT=$s timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu > gpurun_out/r03/bench_split_$s.json 2> /dev/null
  echo "split=$s" $(python - <<PY
import json
d=json.load(open("gpurun_out/r03/bench_split_$s.json"))
print(round(d["value"]), round(d["ms_per_step"]*1e3,1), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"]*1e3,1), round(d["roofline"]["ms_per_launch"]*1e3,1), d["clocks"]["sm_mhz"])
PY
)
done
