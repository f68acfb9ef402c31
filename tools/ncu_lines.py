"""per-source-line share of warp samples and executed instructions of one kernel of an .ncu-rep (captured with
--import-source on, code built with -lineinfo); read on the CPU box:  python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
agg, cur = collections.OrderedDict(), None
for r in rows[hi + 1:]:
    if len(r) < len(h) - 2:
        continue
    if r[0].strip():
        cur = (int(r[0]), r[1].strip()[:95])
    try:
        s, e = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0, collections.Counter()])
    a[0] += s
    a[1] += e
    for i in stall_cols:
        try:
            a[2][h[i][6:]] += int(r[i])
        except ValueError:
            pass
tot, tote = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"# {rep}: {tot} warp samples, {tote} warp instructions executed; per source line of tile_kernel.cuh: samples, share, share of executed instructions, top stall reasons")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:7d} {100 * a[0] / tot:5.1f}%  exec {100 * a[1] / tote:5.1f}%  L{k[0]}: {k[1]}  [{','.join(f'{n}={v}' for n, v in a[2].most_common(3))}]")
