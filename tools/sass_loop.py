"""Opcode histogram of the steady-state march loop of k_chain_fused in a built library (static count over the five unrolled
phases, i.e. per 5 march steps): sass_loop.py <lib.so> <substring of the mangled kernel name>.
The loop is taken as the code between the first mbarrier try-wait and the last backward uniform branch before EXIT."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
cur = None; body = []
for ln in out:
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: body.append((int(m.group(1), 16), m.group(2).strip()))
first = next(i for i, (a, t) in enumerate(body) if "TRYWAIT" in t)
last = max(i for i, (a, t) in enumerate(body) if t.startswith("BRA.U") and i > first and int(t.split("0x")[-1], 16) <= body[first][0] + 0x200 and i < len(body) - 1 and "EXIT" not in t and body[i][0] < body[-1][0])
# restrict to the first backward branch that closes the chunk loop (before EXIT)
ex = next(i for i, (a, t) in enumerate(body) if t.startswith("EXIT") and i > first)
last = max(i for i in range(first, ex) if body[i][1].startswith("BRA"))
h = collections.Counter()
for a, t in body[first:last + 1]:
    op = t.split()
    o = op[1] if op[0].startswith("@") else op[0]
    h[o.split(".")[0]] += 1
n = sum(h.values())
print(f"loop {body[first][0]:#x}..{body[last][0]:#x}: {n} instructions / 5 steps = {n/5:.1f} per step, {16*n/1024:.1f} KB")
print("  " + " ".join(f"{o}:{c/5:.1f}" for o, c in h.most_common(40)))
