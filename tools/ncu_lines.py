#!/usr/bin/env python
"""Executed warp instructions per CUDA source line of one kernel: joins `ncu --page source --csv`
(per-SASS-instruction counters, address order) with `nvdisasm -g` (line of every instruction).
usage: ncu_lines.py <src.csv> <kernel.sass> <kernel-name-substring> [tokens]"""
import csv, re, sys, collections
src, sass, pat = sys.argv[1:4]
tokens = float(sys.argv[4]) if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src)))
kern, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; kern.append(cur)
    elif r and r[0] == 'Address': cur['hdr'] = r
    elif r and cur is not None and len(r) > 5: cur['rows'].append(r)
k = [x for x in kern if pat in x['name']][-1]
h = k['hdr']; ie = h.index('Instructions Executed'); ss = h.index('Warp Stall Sampling (All Samples)')
lines = []; line = ('?', 0)
for l in open(sass):
    m = re.search(r'File "([^"]+)", line (\d+)', l)
    if '//## File' in l and m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l): lines.append((line, l.strip()))
assert len(lines) == len(k['rows']), (len(lines), len(k['rows']))
per = collections.Counter(); st = collections.Counter()
for (ln, txt), r in zip(lines, k['rows']):
    per[ln] += int(r[ie]); st[ln] += int(r[ss])
tot = sum(per.values()); stt = sum(st.values())
print(k['name'], 'total inst', tot, 'per token', tot / tokens if tokens else '')
for ln, n in sorted(per.items()):
    if n * 200 > tot or st[ln] * 100 > stt:
        try: text = open('ldagibbssampling_b200/csrc/' + ln[0]).read().split('\n')[ln[1] - 1].strip()[:90]
        except Exception: text = ''
        print(f'{ln[0][:14]:14s}{ln[1]:5d} {100*n/tot:5.1f}% inst {100*st[ln]/stt:5.1f}% stall' + (f' {n/tokens:6.1f}/tok' if tokens else '') + '  ' + text)
