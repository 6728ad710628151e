"""Executed instructions and stall samples of one kernel per CUDA SOURCE LINE: joins the SASS rows of an ncu report
(`ncu -i rep --page source --csv`, rows in program order) with the line table of the same kernel in the library
(`nvdisasm -g -c` of the cubin extracted with cuobjdump).  usage: tools/ncu_by_line.py rep.ncu-rep <mangled kernel name> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "tumorgrowthtoolkit_b200", "csrc", "libtgtk_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(kern + ":"))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, data = rows[1], rows[2:]
iN, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
why = collections.defaultdict(collections.Counter)
assert len(data) == len(lines), (len(data), len(lines))
cnt, smp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r, ln in zip(data, lines):
    cnt[ln] += int(r[iN]); smp[ln] += int(r[iS])
    for i, h in stall_cols:
        why[ln][h] += int(r[i] or 0)
    src = r[iSrc].split()
    ops[ln][(src[1] if src[0].startswith("@") else src[0]).split(".")[0]] += int(r[iN])
tot, ts = sum(cnt.values()), sum(smp.values())
print(f"{kern}: {tot} warp instructions, {ts} samples")
srcs = {}
order = sorted(cnt, key=lambda k: -(cnt[k] / tot + smp[k] / max(ts, 1)))[:top]
for (f, n) in order:
    c = cnt[(f, n)]
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(root, "tumorgrowthtoolkit_b200", "csrc", f)).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
    mix = " ".join(f"{o}:{100 * v // max(c, 1)}" for o, v in ops[(f, n)].most_common(4))
    reasons = " ".join(f"{h}:{v}" for h, v in why[(f, n)].most_common(3) if v)
    print(f"{100 * c / tot:5.1f}% instr {100 * smp[(f, n)] / max(ts, 1):5.1f}% stall | {f}:{n} | {text} | {mix} | {reasons}")
