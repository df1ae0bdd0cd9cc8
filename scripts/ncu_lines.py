"""Join `nvdisasm -g` line info with an `ncu --page source --csv` SASS dump: dynamic warp-instruction
counts per source line.  usage: ncu_lines.py dis_g.txt src.csv mangled_kernel_name [top]"""
import csv, re, sys, collections
dis, src, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':'))
cur = ('?', 0); seq = []
for l in lines[start + 1:]:
    if l.startswith('//---') or '.section' in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): seq.append(cur)
rows = list(csv.reader(open(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > idx['Instructions Executed'] and r[idx['Instructions Executed']].isdigit()]
data = data[:len(seq)]
assert len(data) == len(seq), (len(data), len(seq))
agg = collections.Counter(); st = collections.Counter(); samp = collections.Counter()
tot = 0
for (f, ln), r in zip(seq, data):
    e = int(r[idx['Instructions Executed']]); agg[(f, ln)] += e; st[(f, ln)] += 1; tot += e
    samp[(f, ln)] += int(r[idx['# Samples']])
print('total dynamic', tot, 'static', len(seq))
src_cache = {}
def text(f, ln):
    import glob
    if f not in src_cache:
        c = glob.glob('/root/repo/rududu_image_codec_b200/csrc/' + f)
        src_cache[f] = open(c[0]).read().split('\n') if c else []
    s = src_cache[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ''
for (f, ln), e in agg.most_common(top):
    print('%5.1f%% dyn %4d static %6d samp  %s:%d  %s' % (100 * e / tot, st[(f, ln)], samp[(f, ln)], f, ln, text(f, ln)))
