#!/bin/bash
# quick GPU check of the volume decoder: parity test, A/B timing against cuDNN, ncu launch list, train-step bench with both
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "vol_decoder" 2>&1 | grep "parity\] vol_decoder \(TF32\|vs\)\|passed\|failed\|Error" | cut -c1-220
python scripts/time_decoder.py hn tf32 | tail -1; python scripts/time_decoder.py cudnn tf32 | tail -1
python scripts/time_decoder.py hn fp32 | tail -1; python scripts/time_decoder.py cudnn fp32 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dec_launches.csv python scripts/time_decoder.py hn tf32 > /dev/null 2>&1
python - <<P
import csv
rows = list(csv.reader(open("gpurun_out/dec_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
rows = rows[hdr + 1:]
names = [(r[4], float(r[-1].replace(",", ""))) for r in rows if len(r) > 6]
n = len(names) // 4
tot = 0
for nm, t in names[-n:]:
    tot += t
    print(f"{t/1000:9.1f} us  {nm[:70]}")
print("total", round(tot / 1000, 1), "us")
P
bash scripts/gpu_exp.sh HN_DECODER=hn HN_DECODER=cudnn
