"""Stall-reason breakdown of an ncu SASS page per ROLE of a warp-specialised kernel (roles = source line ranges).
    python tools/ncu_roles.py sass.csv disasm.txt kernel_prefix source.cu name:first:last [name:first:last ...]"""
import csv, re, sys
from collections import defaultdict
sass_csv, dis, prefix, src_path = sys.argv[1:5]
roles = [(a.split(":")[0], int(a.split(":")[1]), int(a.split(":")[2])) for a in sys.argv[5:]]
base = src_path.split("/")[-1]
txt = open(dis).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = [i for i, l in enumerate(txt) if i > start and l.startswith("//---------------------")]
end = end[0] if end else len(txt)
cur, ins = None, []
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))] + [(f.split("/")[-1], int(n)) for f, n in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        cur = next((c[1] for c in chain if c[0] == base), None)
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv))); hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(float))
for (s, line), r in zip(ins, data):
    role = next((n for n, a, b in roles if line and a <= line <= b), "other")
    agg[role]["inst"] += int(r[ix["Instructions Executed"]]); agg[role]["samples"] += int(r[ix["# Samples"]])
    for h in stalls:
        agg[role][h] += int(r[ix[h]] or 0)
tot = sum(v["inst"] for v in agg.values()); tots = sum(v["samples"] for v in agg.values())
for role, v in agg.items():
    top = sorted(((h, v[h]) for h in stalls), key=lambda kv: -kv[1])[:6]
    print(f"{role:6s} inst {v['inst'] / tot * 100:5.1f}%  samples {v['samples'] / tots * 100:5.1f}%  " + "  ".join(f"{h[6:]} {x / max(v['samples'], 1) * 100:.0f}%" for h, x in top))
