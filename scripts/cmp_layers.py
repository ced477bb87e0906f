import json, sys
a = {r["name"]: r for r in json.load(open(sys.argv[1]))["rows"]}
b = {r["name"]: r for r in json.load(open(sys.argv[2]))["rows"]}
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
print("total", round(sum(r["ms"] for r in a.values()), 3), "vs", round(sum(r["ms"] for r in b.values()), 3))
for r in sorted(a.values(), key=lambda r: -r["ms"])[:n]:
    o = b.get(r["name"], {"ms": 0})
    print(f"{r['name']:32s} A {r['ms']:.3f}  B {o['ms']:.3f}  bn {r.get('block_n')} {round(r.get('gbs') or 0)} / {round(o.get('gbs') or 0)} GB/s")
