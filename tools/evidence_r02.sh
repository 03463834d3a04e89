#!/bin/bash
# Round-2 evidence run on the GPU box: headline bench line, ncu launch list, one ncu --set full capture of k_fit,
# supplementary workloads.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python bench.py > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err; tail -c 400 $O/r02_bench_c5.json
CMD="python bench.py --genes 1184 --steps 2 --warmup 3 --no-cpu"
timeout 200 $CMD > $O/r02_bench_small.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv $CMD > $O/r02_ncu_launches.log 2>&1
CMD2="python bench.py --genes 148 --steps 1 --warmup 1 --no-cpu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o $O/r02_prof_fit $CMD2 > $O/r02_ncu_fit.log 2>&1; tail -2 $O/r02_ncu_fit.log | cut -c1-160
for W in "c1 2368" "c3 296" "c4 8" "c5-n10k 296" "c2 152" "c5-m256 148"; do
  set -- $W
  timeout 240 python bench.py --workload $1 --genes $2 --steps 2 --warmup 3 --no-cpu > $O/r02_bench_$1.json 2> $O/r02_bench_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r02_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$1 failed", e)
PY
done
