"""Per-region dynamic instruction breakdown of generate_slots_kernel from an ncu source CSV + nvdisasm dump."""
import csv, re, sys
from collections import defaultdict
ncu_csv, sass, kname, npairs = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
dyn = [(int(r[ci]), int(r[cs])) for r in rows[2:] if len(r) > ci]
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text." + kname))
cur = None; stat = []
for l in lines[start + 1:]:
    if l.startswith("//--------------------- "): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: stat.append((cur, m.group(2)))
L = open('simuscop_b200/csrc/gen_fast.cu').read().split('\n')
def find(s, start=0):
    for i in range(start, len(L)):
        if s in L[i]: return i + 1
    raise Exception(s)
k0 = find('generate_slots_kernel(const GenParams P)')
pa = find('phase A: one Philox')
marks = [('philox fns', (find('mulwide(uint32_t a'), find('f_draw_pos(uint32_t u'))), ('coop_lookup', (find('int coop_lookup('), find('int uni_lookup('))),
 ('uni_lookup', (find('int uni_lookup('), find('int f_ndigits('))), ('ndigits', (find('int f_ndigits('), find('c_pow10[10]'))),
 ('call_base(slow)', (find('uint32_t call_base('), find('uint32_t window_code('))), ('window_code(slow)', (find('uint32_t window_code('), find('int slow_read('))),
 ('slow_read', (find('int slow_read('), find('// quality lookup: row of QP'))), ('qual_lookup', (find('// quality lookup: row of QP'), k0)),
 ('prologue', (k0, find('consecutive pairs go to consecutive warps'))), ('bin+frag', (find('consecutive pairs go to consecutive warps'), find('prefetch the packed windows'))),
 ('prefetch', (find('prefetch the packed windows'), find('header digits'))), ('header', (find('header digits'), find('uint32_t lens = 0;'))),
 ('mate setup', (find('uint32_t lens = 0;'), pa)), ('phaseA', (pa, find('// ---- header', pa))), ('hdr store', (find('// ---- header', pa), find('phase C, fast path'))),
 ('phaseC fast', (find('phase C, fast path'), find('s_xsave[c * 32 + lane] = x2[c]') - 2)), ('tail', (find('s_xsave[c * 32 + lane] = x2[c]') - 2, find('pass 2: scan of the record')))]
agg = defaultdict(lambda: [0, 0, 0]); tot = 0
for (loc, txt), (n, smp) in zip(stat, dyn):
    tot += n; key = 'other'
    if loc is None: key = 'noloc'
    elif loc[0] != 'gen_fast.cu': key = loc[0]
    else:
        for name, r in marks:
            if r[0] <= loc[1] < r[1]: key = name; break
    agg[key][0] += 1; agg[key][1] += n; agg[key][2] += smp
tots = sum(v[2] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-28s static %5d dyn %5.1f%%  per pair %7.1f  samples %5.1f%%' % (k, v[0], 100 * v[1] / tot, v[1] / npairs, 100 * v[2] / tots))
print('total per pair %.1f, static %d' % (tot / npairs, len(stat)))
