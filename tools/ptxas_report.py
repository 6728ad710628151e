"""Registers / spills / shared memory of every kernel from `nvcc -Xptxas -v` output (the Makefile passes it):
    make -C tumorgrowthtoolkit_b200/csrc -B 2>&1 | python tools/ptxas_report.py [filter-substring ...]"""
import re, subprocess, sys
txt = sys.stdin.read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, "
                  r"(\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers, used \d+ barriers"
                  r"(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt)
names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
for n, e in zip(names, ents):
    n2 = re.sub(r"\(.*", "", n).replace("void tgtk::", "")
    if len(sys.argv) > 1 and not any(f in n2 for f in sys.argv[1:]):
        continue
    print(f"{n2:58s} regs {e[4]:>4s}  spill st/ld {e[2]}/{e[3]}  smem {e[6] or 0}")
for l in txt.splitlines():
    if "error" in l or "warning" in l:
        print(l)
