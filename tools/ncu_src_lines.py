"""Per-source-line instruction and stall-sample shares of one kernel in an ncu report captured with --import-source on.

    python tools/ncu_src_lines.py report.ncu-rep [top_n]
"""
import csv, subprocess, sys
rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, per, stalls = None, {}, {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        ie, isamp, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and r and r[0].isdigit():
        try:
            per[int(r[0])] = (int(r[ie]), int(r[isamp]), int(r[te]), r[1])
        except ValueError:
            continue
        d = {}
        for i in st:
            try:
                v = int(r[i])
            except ValueError:
                continue
            if v and v < 10 ** 7:
                d[hdr[i][6:]] = v
        stalls[int(r[0])] = d
tot, ts = sum(v[0] for v in per.values()), sum(v[1] for v in per.values())
print("warp instructions", tot, "samples", ts)
agg = {}
for d in stalls.values():
    for k, v in d.items():
        agg[k] = agg.get(k, 0) + v
print("stall samples:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for ln, (n, s, t, src) in sorted(per.items(), key=lambda kv: -kv[1][1])[:topn]:
    top = sorted(stalls[ln].items(), key=lambda kv: -kv[1])[:3]
    print(f"{ln:4d} {n / tot * 100:5.1f}%i {s / ts * 100:5.1f}%s thr/inst {t / max(n, 1):4.1f} {src.strip()[:90]}  {top}")
