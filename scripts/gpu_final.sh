#!/bin/bash
# End-of-round evidence: full GPU test suite + smoke + default bench, launch list of the volume decoder alone.
# (The ncu captures are a second call, scripts/gpu_profile.sh TAG: gpurun brings back at most 64 MiB per call.)
TAG=${1:-r02b}
bash scripts/gpu_round2.sh
cp gpurun_out/bench.log gpurun_out/bench_${TAG}_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dec_launches.csv python scripts/time_decoder.py hn tf32 > /dev/null 2>&1
python - <<P > gpurun_out/${TAG}_decoder_launches.txt
import csv
rows = list(csv.reader(open("gpurun_out/dec_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
rows = rows[hdr + 1:]
names = [(r[4], float(r[-1].replace(",", ""))) for r in rows if len(r) > 6]
n = len(names) // 4
tot = 0
print("# ncu launch list (gpu__time_duration.sum, --clock-control none) of one forward + backward of the volume decoder,")
print("# python scripts/time_decoder.py hn tf32 (4th repetition); cold-cache, serialised times")
for nm, t in names[-n:]:
    tot += t
    print(f"{t/1000:9.1f} us  {nm[:110]}")
print("total", round(tot / 1000, 1), "us")
P
tail -3 gpurun_out/${TAG}_decoder_launches.txt
