"""Top SASS lines of an ncu report by stall samples (with the source file:line when -lineinfo + --import-source were used)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
print("columns:", [h for h in hdr][:40])
isrc = hdr.index("Source")
isamp = next(i for i, h in enumerate(hdr) if h.startswith("Warp Stall Sampling (All"))
iex = hdr.index("Instructions Executed")
body = []
for k, r in enumerate(rows[2:]):
    try: body.append((int(r[isamp]), int(r[iex]), k, r[isrc]))
    except Exception: pass
tot = sum(b[0] for b in body)
print("total samples", tot)
if len(sys.argv) > 3:   # full table (index, samples, executed, SASS) for source mapping with nvdisasm -g
    with open(sys.argv[3], "w") as f:
        for s_, ex, k, t in sorted(body, key=lambda b: b[2]):
            f.write("%d\t%d\t%d\t%s\n" % (k, s_, ex, t.strip()))
for s, ex, k, t in sorted(body, reverse=True)[:n]:
    print("%6.2f%%  ex=%-9d #%-5d %s" % (100.0 * s / max(tot, 1), ex, k, t.strip()[:110]))
