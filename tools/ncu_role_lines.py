"""Per source line of ONE role (line range) of a warp-specialised kernel: share of the role's samples and top stall reasons.
    python tools/ncu_role_lines.py sass.csv disasm_gi.txt kernel_prefix source.cu first last [top]"""
import csv, re, sys
from collections import defaultdict
sass_csv, dis, prefix, src_path = sys.argv[1:5]
lo, hi = int(sys.argv[5]), int(sys.argv[6]); top = int(sys.argv[7]) if len(sys.argv) > 7 else 25
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
        cur = (next((c[1] for c in chain if c[0] == base), None), chain[0])
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv))); hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(float))
for (s, c), r in zip(ins, data):
    if not c or not c[0] or not (lo <= c[0] <= hi):
        continue
    key = (c[0], c[1] if c[1][0] != base else None)
    agg[key]["inst"] += int(r[ix["Instructions Executed"]]); agg[key]["samples"] += int(r[ix["# Samples"]])
    for h in stalls:
        agg[key][h] += int(r[ix[h]] or 0)
tot = sum(v["inst"] for v in agg.values()); tots = sum(v["samples"] for v in agg.values())
src = open(src_path).read().split("\n")
print(f"role lines {lo}-{hi}: {tot:.0f} warp instructions, {tots:.0f} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((h, v[h]) for h in stalls), key=lambda kv: -kv[1])[:3]
    where = src[k[0] - 1].strip()[:70] + (f"  <{k[1][0]}:{k[1][1]}>" if k[1] else "")
    print(f"{v['samples'] / tots * 100:5.1f}% samp {v['inst'] / tot * 100:5.1f}% inst  {k[0]:4d} " + " ".join(f"{h[6:]}={x / max(v['samples'], 1) * 100:.0f}%" for h, x in st) + "  | " + where)
