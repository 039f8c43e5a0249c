"""Cluster the mbarrier polling instructions of an ncu SASS page by call site (see tools/ncu_lines.py for the inputs).
    python tools/ncu_spin.py sass.csv disasm.txt kernel_prefix source.cu wait_first_line wait_last_line body_first_line"""
import csv, re, sys
sass_csv, dis, prefix, src_path = sys.argv[1:5]
w0, w1, body0 = int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
base = src_path.split("/")[-1]
txt = open(dis).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = [i for i, l in enumerate(txt) if i > start and l.startswith("//---------------------")]
end = end[0] if end else len(txt)
cur, ins = None, []
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv))); hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
n = min(len(ins), len(data)); tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
src = open(src_path).read().split("\n")
inw = lambda c: c and c[0] == base and w0 <= c[1] <= w1
body = lambda c: c and c[0] == base and c[1] >= body0
i = 0
while i < n:
    if inw(ins[i][1]):
        j = i
        while j < n and (inw(ins[j][1]) or (ins[j][1] and ins[j][1][0] != base)): j += 1
        a = sum(int(data[k][ia]) for k in range(i, j)); s = sum(int(data[k][isamp]) for k in range(i, j))
        prv = next((ins[k][1][1] for k in range(i - 1, -1, -1) if body(ins[k][1])), 0)
        nxt = next((ins[k][1][1] for k in range(j, n) if body(ins[k][1])), 0)
        if a / tot > 0.002 or s / tots > 0.002:
            print(f"{a / tot * 100:5.2f}% inst {s / tots * 100:5.2f}% samp  n={j - i:3d}  after line {prv}: {src[prv - 1].strip()[:60]}  | before {nxt}")
        i = j
    else:
        i += 1
