"""Per-kernel SASS census of the built library: python tools/sass_summary.py > profiles/rN_sass_summary.txt
Counts the mnemonics that show which hardware paths a kernel uses (TMA tensor / bulk copies, mbarrier, fp64 add / mul / fma, shared and
global atomics, warp votes / shuffles / reductions, barriers). cuobjdump runs on the host: no GPU needed."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scouting-ipp_b200", "libsipp.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTMALDG", "UBLKCP", "SYNCS", "UTCHMMA|UTCQMMA|UTCIMMA|UTCOMMA", "HMMA|IMMA|DMMA", "DADD", "DMUL", "DFMA", "DSETP|DMNMX", "ATOMS", "ATOMG|RED", "SHFL", "VOTE", "REDUX",
        "BAR", "LDGSTS", "LDG", "STG", "LDS", "STS", "NANOSLEEP"]
fn, counts, order, total = None, {}, [], collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn)
        fn = re.sub(r"\(.*", "", fn)
        counts[fn] = collections.Counter()
        order.append(fn)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        for w in WANT:
            if re.match("(?:%s)(?:\\.|$)" % w, op):
                counts[fn][w] += 1
print("# SASS census of %s (sm_100a), instruction counts per kernel; blank = 0" % os.path.relpath(lib, ROOT))
print("# no tensor-core mnemonic (UTC*MMA / *MMA) is expected anywhere: nothing on this path is a contraction")
cols = [w.split("|")[0] + ("*" if "|" in w else "") for w in WANT]
print("%-44s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
for fn in order:
    print("%-44s %6d " % (fn[-44:], total[fn]) + " ".join("%7s" % (counts[fn][w] or "") for w in WANT))
