"""Stall samples of an ncu report aggregated per CUDA source line (needs -lineinfo and --import-source on)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several files may be listed one after another, each with its own header row
tot = 0; agg = []
hdr = None; fname = ""
for r in rows:
    if not r: continue
    if len(r) == 1 or (len(r) >= 2 and r[0].startswith("File")):
        fname = r[-1] if len(r) > 1 else r[0]; continue
    if "Source" in r and any(h.startswith("Warp Stall Sampling (All") for h in r):
        hdr = r; isrc = hdr.index("Source"); isamp = next(i for i, h in enumerate(hdr) if h.startswith("Warp Stall Sampling (All")); iln = 0
        iex = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        continue
    if hdr is None: continue
    try:
        sm = int(r[isamp])
    except Exception:
        continue
    ex = r[iex] if iex is not None else ""
    tot += sm
    if sm: agg.append((sm, r[iln], ex, r[isrc].strip()[:120], fname[-40:]))
print("total samples", tot)
for sm, ln, ex, src, f in sorted(agg, reverse=True)[:n]:
    print("%6.2f%%  L%-5s ex=%-9s %s   [%s]" % (100.0 * sm / max(tot, 1), ln, ex, src, f))
