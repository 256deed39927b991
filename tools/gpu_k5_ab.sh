#!/bin/bash
# K5 check: parity tests, then probe timings (k5_ms = non-SCC, k5_scc_diag_ms, k5_scc_full_ms).
TAG=${1:-r2}
mkdir -p gpurun_out
echo "== pytest test_gpu_parity"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scc_end_to_end.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_k5_$TAG.log
for nb in ${K5NBS:-2 3}; do
echo "== NB $nb"
DFTBP_B200_K5_NB=$nb timeout 300 python tools/gpu_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/probe_k5_${nb}_$TAG.log
DFTBP_B200_K5_NB=$nb timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"k5_|ps_" --csv --log-file gpurun_out/k5_times_$TAG.csv python tools/gpu_probe.py fec4096 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/k5_times_$TAG.csv")) if len(r)>10]
hdr=rows[0]; d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[hdr.index("Kernel Name")][:60]].append(float(r[-1]))
for n,v in d.items(): print("dur",n,len(v),"launches, median ns",sorted(v)[len(v)//2],"min",min(v), "max", max(v))
PY
done
