#!/bin/bash
# DRAM traffic of the resident-weight conv kernel variants (development aid): microbench_conv.py under
# `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`, one line per launch.
cd "$(dirname "$0")/.."
for dbg in "$@"; do
  echo "== AV1P_CR_DEBUG=$dbg"
  AV1P_CR_DEBUG=$dbg ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_res -c 28 --csv python tools/microbench_conv.py 2>/dev/null \
    | python -c "
import csv, sys, collections
rows = [r for r in csv.reader(sys.stdin) if len(r) > 14 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {'name': r[4]})[r[12]] = (float(r[14].replace(',', '')), r[13])
for i, e in by.items():
    print(i, e['name'][:60], {k: v for k, v in e.items() if k != 'name'})
" | tail -8
done
