"""SASS around the stage release of k_pipe<OP_TIME_AVG> for a library build: the shared-memory loads of the stage, the release
sequence, and per-kernel counts of the Blackwell-native instructions (UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = bulk copies,
SYNCS = mbarrier ops, LDGSTS = cp.async).  usage: sass_release_excerpt.py <libtse_cuda.so>"""
import re, subprocess, sys, collections
lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", sass)[1:]
pat = re.compile(r"/\*([0-9a-f]{4})\*/\s+(.*?);")
print("%-34s %8s %8s %7s %6s %7s %6s" % ("kernel", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "regs"))
for f in funcs:
    name = f.split("\n", 1)[0]
    short = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void tse::", "").replace("tse::", "")
    c = collections.Counter()
    for m in pat.finditer(f):
        op = m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1]
        c[op.split(".")[0]] += 1
    if c["UTMALDG"] or c["SYNCS"] or "remap" in short or "nbr" in short:
        print("%-34s %8d %8d %7d %6d %7d" % (short[:34], c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["SYNCS"], c["LDGSTS"]))
for f in funcs:
    if "k_pipeILi5E" not in f.split("\n", 1)[0]:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in pat.finditer(f)]
    # the consumer loop: first SYNCS.PHASECHK after the last BAR.SYNC ... up to the arrive that follows the LDS burst
    idx = [i for i, (a, s) in enumerate(ins) if "SYNCS.ARRIVE" in s and "A1T0" in s or "SYNCS.ARRIVE" in s and "ART0" in s]
    k = idx[-1]
    j = k
    while j > 0 and "TRYWAIT" not in ins[j][1]:
        j -= 1
    print("\nk_pipe<OP_TIME_AVG>, consumer: wait on the full barrier -> loads of the stage -> release (memory and barrier instructions only)")
    for a, s in ins[j:k + 1]:
        if re.search(r"SYNCS|LDS|MEMBAR|FENCE|BAR\.|WARPSYNC|DEPBAR", s) and "@!PT" not in s:
            print("  /*%04x*/  %s" % (a, s))
