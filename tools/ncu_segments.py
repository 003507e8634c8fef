"""Split the warp-stall samples of an `ncu --set full --import-source on` capture into the code segments between
barriers / mbarrier waits of a kernel (SASS order), with the MUFU mix of each segment to recognise the phase.
  python tools/ncu_segments.py gpurun_out/x.ncu-rep [min_share]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = {h: hdr.index(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
tot = sum(int(r[isamp] or 0) for r in data)
print(rows[0][1][:80], "| total samples", tot, "| SASS instructions", len(data))
new = lambda i: {"start": i, "samples": 0, "n": 0, "mufu": {}, "exec": 0, "stalls": {}, "ldg": 0, "tmem": 0}
seg, cur = [], new(0)
for i, r in enumerate(data):
    s, n, ex = r[isrc], int(r[isamp] or 0), int(r[iexec] or 0)
    cur["samples"] += n; cur["n"] += 1; cur["exec"] += ex
    for k, c in stall_cols.items():
        v = int(r[c] or 0)
        if v: cur["stalls"][k] = cur["stalls"].get(k, 0) + v
    if "MUFU" in s:
        op = s.split("MUFU.")[1].split()[0]
        cur["mufu"][op] = cur["mufu"].get(op, 0) + ex
    if "LDG" in s: cur["ldg"] += ex
    if "LDTM" in s or "STTM" in s: cur["tmem"] += ex
    if "BAR.SYNC" in s or "BAR.RED" in s or "SYNCS.PHASECHK" in s:
        cur["end"], cur["why"] = i, s.strip()[:34]; seg.append(cur); cur = new(i + 1)
cur["end"], cur["why"] = len(data) - 1, "end"; seg.append(cur)
for sg in seg:
    if sg["samples"] < min_share * tot: continue
    top = sorted(sg["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print(f'{sg["start"]:5d}-{sg["end"]:5d} samples {100 * sg["samples"] / tot:5.1f}% exec {sg["exec"] / 1e6:7.1f}M ldg {sg["ldg"] / 1e6:5.1f}M tmem {sg["tmem"] / 1e6:5.1f}M '
          f'mufu {dict((k, round(v / 1e6, 1)) for k, v in sg["mufu"].items())} | {[(k[6:], round(100 * v / max(sg["samples"], 1))) for k, v in top]} | {sg["why"]}')
