"""Per-source-line instruction and stall-sample shares of one kernel: joins `ncu --page source --print-source sass --csv`
with the line table of `nvdisasm -g -c` (run both here, no GPU needed).  usage: ncu_lines.py sass.csv dis.txt kernel [top]"""
import csv, re, sys
sass, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
off2line, cur = {}, None
for l in lines[start + 1:]:
    if l.startswith('.text.'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass)))
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
base, agg, tots, toti = None, {}, 0, 0
for r in rows[hi + 1:]:
    try: a = int(r[0], 16); s = float(r[ci['# Samples']]); ie = float(r[ci['Instructions Executed']])
    except Exception: continue
    if base is None: base = a
    ln = off2line.get(a - base)
    d = agg.setdefault(ln, [0, 0, {}]); d[0] += s; d[1] += ie; tots += s; toti += ie
    for st in stalls:
        v = float(r[ci[st]] or 0)
        if v: d[2][st] = d[2].get(st, 0) + v
src = {}
import os
for ln in agg:
    if ln and ln[0] not in src:
        for root in ('/root/repo/cgraphs_b200/csrc/',):
            if os.path.exists(root + ln[0]): src[ln[0]] = open(root + ln[0]).read().split('\n')
print("samples %d  warp instructions %d" % (tots, toti))
for ln, (s, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src.get(ln[0], [''] * 100000)[ln[1] - 1].strip()[:90] if ln and ln[0] in src else ''
    t3 = ' '.join('%s=%d%%' % (k[6:], 100 * v / max(s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print("%5.1f%%smp %5.1f%%ins %-22s| %-90s | %s" % (100 * s / tots, 100 * ie / toti, '%s:%d' % ln if ln else '?', text, t3))
