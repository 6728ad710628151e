"""For every global load inside a kernel's SASS: how many instructions later is its destination first read?
    cuobjdump -sass lib.so | python tools/sass_first_use.py <mangled-name-substring>"""
import re, sys
pat = sys.argv[1]
lines, on = [], False
for l in sys.stdin:
    if "Function :" in l:
        on = pat in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
regs = lambda s: set(re.findall(r"\bR(\d+)\b", s))
out = []
for i, (addr, ins) in enumerate(lines):
    m = re.match(r"(@!?U?P\d+\s+)?LDG\.E\.64\s+R(\d+),", ins)
    if not m:
        continue
    d = int(m.group(2))
    dst = {str(d), str(d + 1)}
    for j in range(i + 1, len(lines)):
        body = lines[j][1]
        ops = body.split(",", 1)[1] if "," in body else ""
        if body.startswith("ST") or "STG" in body or "STS" in body:
            ops = body
        if regs(ops) & dst:
            nbar = sum(1 for k in range(i, j) if "BAR.SYNC" in lines[k][1])
            out.append((addr, ins[:44], j - i, nbar, lines[j][1][:50]))
            break
for o in out:
    print(f"{o[0]:05x} {o[1]:46s} first use +{o[2]:4d} instr, {o[3]} BAR between: {o[4]}")
