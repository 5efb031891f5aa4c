import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None; agg = collections.OrderedDict(); n = 0
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    n += 1
    if n <= skip: continue
    name = r[4].split('(')[0].replace('void ', '').replace('nfs::<unnamed>::', '')
    agg.setdefault(name, []).append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
print("total %.1f us over %d launches" % (tot / 1e3, sum(len(v) for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-64s n=%3d  avg %8.1f us  sum %8.1f us  share %5.1f%%" % (k[:64], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
