"""Static listing of the built library (no GPU needed): per kernel the registers / stack / static shared memory that
`cuobjdump -res-usage` reports and a histogram of the SASS mnemonics from `cuobjdump -sass`, grouped so that the memory
path is visible (LDG/STG widths, shared-memory and global atomics, bulk copies, tensor-core ops, barriers).

usage: python profiles/scripts/sass_summary.py [path/to/libdet_b200.so] > profiles/r02/sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "object-detection-pytorch-rust_b200", "det_b200", "_lib", "libdet_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)  # argument list
    return name.replace("det::", "")


res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv)
        cur = None

sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
hist = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1

names = demangle(sorted(set(usage) | set(hist)))
groups = [("LDG.128", r"^LDG\.E.*\.128"), ("LDG.64", r"^LDG\.E.*\.64"), ("LDG other", r"^LDG"), ("STG.128", r"^STG\.E.*\.128"),
          ("STG.64", r"^STG\.E.*\.64"), ("STG other", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("ATOMS", r"^ATOMS"),
          ("ATOMG/RED", r"^(ATOMG|RED|ATOM\b)"), ("SHFL", r"^SHFL"), ("VOTE/MATCH", r"^(VOTE|MATCH)"), ("BAR", r"^BAR"),
          ("MUFU", r"^MUFU"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)"), ("UBLKCP (bulk copy)", r"^UBLKCP"),
          ("UTMA* (tensor copy)", r"^UTMA"), ("UTC*MMA (tcgen05)", r"^UTC"), ("HMMA/IMMA (mma.sync)", r"^(HMMA|IMMA)"),
          ("SYNCS (mbarrier)", r"^SYNCS"), ("ACQBULK/PDL", r"^(ACQBULK|PREEXIT)")]
print(f"# {os.path.basename(so)}: registers / shared memory / SASS mnemonic groups per kernel (cuobjdump, sm_100a)")
print("# No kernel of this path is GEMM-shaped (byte / index / compare work, HBM- or issue-bound), so UTC*MMA = 0 is expected;")
print("# UBLKCP = 0 because the bulk-copy variant of the dense decode measured slower than plain 16-byte loads and was removed")
print("# (profiles/README.md, r01 dense head table).\n")
for mangled in sorted(hist, key=lambda k: short(names[k])):
    h = hist[mangled]
    u = usage.get(mangled, {})
    total = sum(h.values())
    print(f"{short(names[mangled])}")
    print(f"    REG {u.get('REG', '?')}  STACK {u.get('STACK', '?')}  SHARED(static) {u.get('SHARED', '?')}  LOCAL {u.get('LOCAL', '?')}  "
          f"SASS instructions {total}")
    cells = []
    left = collections.Counter(h)
    for label, pat in groups:
        n = sum(c for k, c in left.items() if re.match(pat, k))
        for k in [k for k in left if re.match(pat, k)]:
            del left[k]
        if n:
            cells.append(f"{label} {n}")
    print("    " + ", ".join(cells))
