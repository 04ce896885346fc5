"""Summarise an `ncu --page source --csv` dump: instruction mix, stall samples, hottest SASS lines."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ci, ce, cs = ix['Source'], ix['Instructions Executed'], ix['# Samples']
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
mix, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    try:
        n = float(r[ce])
    except Exception:
        continue
    toks = r[ci].split()
    op = toks[0] if toks else '?'
    if op.startswith('@') and len(toks) > 1:
        op = toks[1]
    op = op.split('.')[0]
    mix[op] += n; tot += n
    s = float(r[cs] or 0)
    samp[op] += s
    for h in stall_cols:
        try: stalls[h] += float(r[ix[h]] or 0)
        except Exception: pass
    lines.append((s, n, r[ix['Address']], r[ci][:90], {h: r[ix[h]] for h in stall_cols if r[ix[h]] not in ('', '0')}))
S = sum(samp.values())
print("total warp instructions executed: %.3e" % tot)
for op, n in mix.most_common(22):
    print("%-12s %6.2f%% of instr   %5.1f%% of samples" % (op, 100 * n / tot, 100 * samp[op] / max(S, 1)))
print("--- stall reasons (all samples)")
T = sum(stalls.values())
for h, v in stalls.most_common(10):
    print("%-28s %5.1f%%" % (h, 100 * v / max(T, 1)))
print("--- hottest lines")
for s, n, addr, src, st in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6.2f%% %s %-90s %s" % (100 * s / max(S, 1), addr[-5:], src, dict(sorted(st.items(), key=lambda kv: -float(kv[1]))[:2])))

for key in ("stall_no_inst", "stall_branch_resolving", "stall_sleep"):
    cols = [h for h in stall_cols if h.startswith(key)]
    if not cols:
        continue
    print("--- lines with the most %s samples" % key)
    ranked = sorted(lines, key=lambda l: -sum(float(l[4].get(c, 0) or 0) for c in cols))[:12]
    for s_, n, addr, src, st in ranked:
        print("%7d  %s %-80s" % (sum(float(st.get(c, 0) or 0) for c in cols), addr[-5:], src))
