"""Per-kernel SASS opcode histogram of the built library (profiles/r2_sass_opcodes.txt):
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt
Shows, per kernel, the counts of the opcodes that prove the Blackwell-native paths (UTCIMMA / UTCBAR /
LDTM = tcgen05 MMA, commit, tensor-memory load; UBLKCP = bulk TMA copy; SYNCS = mbarrier; DFMA/DADD/DMUL
= the fp64 pipe) and the 12 most frequent opcodes."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cpu-gpu-tfhe_b200", "libtfhe_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|tfhe_b200::", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "IMAD", "LDS", "STS", "SHFL",
       "ATOMS", "BAR", "LDG", "STG", "LDL", "STL"]
print("library:", os.path.relpath(lib, ROOT))
for k, h in hist.items():
    print("\n== %s  (%d instructions)" % (k, sum(h.values())))
    print("   " + "  ".join("%s %d" % (o, h[o]) for o in KEY if h[o]))
    print("   top: " + "  ".join("%s %d" % (o, c) for o, c in h.most_common(12)))
