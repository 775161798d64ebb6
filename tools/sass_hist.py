"""Static SASS opcode histogram of one kernel in a built library: sass_hist.py <lib.so> <substring of the mangled name> [top]."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
cur = None; hist = {}
for ln in out:
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur: hist[cur][m.group(1).split(".")[0]] += 1
for k, h in hist.items():
    if pat in k:
        print(k[-90:], "total", sum(h.values()))
        print("  " + " ".join(f"{o}:{c}" for o, c in h.most_common(top)))
