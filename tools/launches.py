"""Print an ncu launch-list csv (gpu__time_duration, dram bytes, warp instructions) as a table."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
d = defaultdict(dict)
for r in rows:
    d[int(r[0])]['name'] = r[4][:44]
    d[int(r[0])][r[12]] = float(r[14].replace(',', ''))
tot = 0.0
for i in sorted(d):
    x = d[i]
    tot += x.get('gpu__time_duration.sum', 0) / 1e3
    print(f"{i:3d} {x['name']:44s} {x.get('gpu__time_duration.sum', 0) / 1e3:8.1f} us  rd {x.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB  "
          f"wr {x.get('dram__bytes_write.sum', 0) / 1e6:6.1f} MB  {x.get('smsp__inst_executed.sum', 0) / 1e6:7.1f} M warp-inst")
print(f'total {tot:.1f} us over {len(d)} launches')
