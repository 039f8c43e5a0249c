"""Attribute the per-instruction counters of an ncu report to CUDA source lines.

ncu's SASS page carries no line column and its CUDA page needs the source at the capture path, so this joins the SASS
page (ncu -i rep --page source --csv --print-source sass) with `nvdisasm -g` of the cubin the report was taken from
(cuobjdump -xelf all build/<file>.o), instruction by instruction.

    python tools/ncu_lines.py sass.csv disasm.txt kernel_mangled_prefix source.cu [top]
"""
import csv, re, sys
from collections import defaultdict


def main():
    sass_csv, dis, prefix, src_path = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    txt = open(dis).read().split("\n")
    start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
    end = [i for i, l in enumerate(txt) if i > start and l.startswith("//---------------------")]
    end = end[0] if end else len(txt)
    cur, ins = None, []
    for l in txt[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            outer = (inl[-1][0].split("/")[-1], int(inl[-1][1])) if inl else None
            cur = ((m.group(1).split("/")[-1], int(m.group(2))), outer)
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2).strip(), cur))
    rows = list(csv.reader(open(sass_csv)))
    hdr, data = rows[1], rows[2:]
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    if len(ins) != len(data):
        print(f"WARNING: {len(ins)} disassembled vs {len(data)} profiled instructions -- different builds?")
    src = open(src_path).read().split("\n")
    base = src_path.split("/")[-1]
    by = defaultdict(lambda: [0, 0])
    for (s, c), r in zip(ins, data):
        key = None
        if c:
            key = c[0] if c[0][0] == base else (c[1] if c[1] and c[1][0] == base else c[0])
        by[key][0] += int(r[ia]); by[key][1] += int(r[isamp])
    tot = sum(v[0] for v in by.values()); tots = sum(v[1] for v in by.values())
    print(f"total warp instructions {tot}, samples {tots}")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        line = src[k[1] - 1].strip()[:100] if k and k[0] == base else str(k)
        print(f"{v[0] / tot * 100:5.2f}% inst {v[1] / tots * 100:5.2f}% samp  {k[1] if k else 0:4d} {line}")


if __name__ == "__main__":
    main()
