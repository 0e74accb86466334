"""Ad-hoc profiler helper: attribute `ncu --page source --csv` SASS-level "Instructions Executed" to source lines using
nvdisasm --print-line-info of the shipped cubin.  usage: sass_lines.py <src.csv> <lines.sass> [top]"""
import csv, re, sys, collections
src, lines = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src)))
h = rows[1]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
ex = []
for r in rows[2:]:
    try:
        ex.append((int(r[ie]), int(r[isamp] or 0), r[h.index("Source")]))
    except Exception:
        pass
# line info: sequence of instructions in order, with current file/line (stop at the next function)
cur = None; order = []
started = False
for ln in open(lines):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        order.append(cur)
    elif ln.startswith("//-------") and ".text." in ln:
        if started:
            break
        started = True
n = min(len(order), len(ex))
agg = collections.Counter(); samp = collections.Counter()
for i in range(n):
    agg[order[i]] += ex[i][0]; samp[order[i]] += ex[i][1]
tot = sum(agg.values()); ts = sum(samp.values())
print("instructions %d, sass %d / lines %d" % (tot, len(ex), len(order)))
for k, v in agg.most_common(top):
    print("%-22s %6d  %5.1f%% inst  %5.1f%% samples" % (k[0], k[1], 100.0 * v / tot, 100.0 * samp[k] / max(1, ts)))
