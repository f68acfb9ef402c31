"""SASS evidence per kernel of libcunqa_b200.so: counts of the mnemonics that prove the data path (TMA bulk-tensor loads /
stores, FP64 FMA, FP64 tensor MMA, mbarrier waits, shuffles).  CPU only:  python tools/sass_extract.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cunqa_b200", "libcunqa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass, architectures {arch}")
want = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DMMA", "FFMA", "HMMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UTCHMMA"]
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        total[cur] += 1
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                counts[cur][w] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':100s} {'instr':>7s} " + " ".join(f"{w:>7s}" for w in want))
for name, dn in zip(counts, demangle):
    short = re.sub(r"\(.*", "", dn).replace("cb200::", "").replace("void ", "")[:100]
    print(f"{short:100s} {total[name]:7d} " + " ".join(f"{counts[name][w]:7d}" for w in want))
