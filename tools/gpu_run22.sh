#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --profile-from-start off --csv --log-file $O/launches_q.csv python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_list.log 2>&1
echo "list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_q.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik]})[r[im]]=r[iv]
for i,v in d.items():
    print(i, v['k'][:60], v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'), v.get('dram__bytes_write.sum'))
PY
