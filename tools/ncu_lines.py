"""Per-source-line and per-opcode dynamic instruction counts from an .ncu-rep (needs -lineinfo)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; cells = float(sys.argv[2]) if len(sys.argv) > 2 else 67108864.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
def page(kind):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", kind], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))
# opcode totals from the SASS-only view (the cuda,sass view repeats an instruction under every line of its inline chain)
rows = page("sass"); hdr = rows[1]; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iA = hdr.index("Source")
byop = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    try: e = int(r[iE])
    except (ValueError, IndexError): continue
    op = r[iA].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0]; byop[o] += e; tot += e
    try: samp[o] += int(r[iS])
    except ValueError: pass
# per-line counts (inclusive of inlined callees attributed to that line)
rows = page("cuda,sass"); hdr = rows[2]; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
agg = []
for r in rows[3:]:
    if len(r) <= iE or r[0] == "": continue
    try: agg.append((int(r[iE]), int(r[0]), r[1], int(r[iS])))
    except ValueError: pass
print("dynamic warp instr", tot, " lane-instr per cell", tot * 32 / cells, " warp-instr per 60-cell step", tot / (cells / 60))
ts = max(1, sum(samp.values()))
print(" ".join(f"{o}:{c*32/cells:.1f}" for o, c in byop.most_common(30)))
agg.sort(reverse=True)
for e, l, s, sm in agg[:top]:
    print(f"{e*32/cells:7.2f}/cell {100*sm/ts:5.1f}%smp  L{l}: {s.strip()[:105]}")
