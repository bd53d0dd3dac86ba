"""Print one step of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else 'pack'
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
seq = [(r[ki], float(r[vi].replace(',', '')) / (1000 if r[ui] == 'ns' else 1)) for r in data if len(r) > vi]
idx = [i for i, (k, v) in enumerate(seq) if marker in k]
a, b = idx[-3], idx[-2]
tot = 0
for k, v in seq[a:b]:
    print(f"{v:10.1f} us  {re.sub(r'[(<].*', '', k.replace('conp::<unnamed>::','').replace('void ',''))[:60]}")
    tot += v
print(f"{tot:10.1f} us  step total ({b-a} launches)")
