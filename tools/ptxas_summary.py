"""Summarise the ptxas -v logs the Makefile leaves in rgbd-pl-slam_b200/build/*.ptxas.log: one row per kernel
(registers, spill bytes, static shared memory, stack)."""
import re, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for path in sorted(glob.glob(os.path.join(ROOT, "rgbd-pl-slam_b200", "build", "*.ptxas.log"))):
    txt = open(path).read().splitlines()
    cur = None
    stack = spill_s = spill_l = 0
    for ln in txt:
        m = re.search(r"Compiling entry function '([^']+)'", ln)
        if m:
            cur = m.group(1); continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and cur:
            stack, spill_s, spill_l = map(int, m.groups()); continue
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", ln)
        if m and cur:
            smem = re.search(r"(\d+) bytes smem", ln)
            rows.append((os.path.basename(path).replace(".ptxas.log", ".cu"), cur, int(m.group(1)), stack, spill_s, spill_l,
                         int(smem.group(1)) if smem else 0))
            cur = None
def short(mangled):
    m = re.search(r"\d(k_[a-z0-9_]+?)(?:E(?:NS|v|P|i)|I(L[ib])(\d+)E)", mangled)
    if not m:
        return re.sub(r"^.*?(k_\w+)$", r"\1", mangled)
    return m.group(1) + ("<%s>" % m.group(3) if m.group(3) else "")
dem = [short(r[1]) for r in rows]
print("| file | kernel | regs | stack B | spill st B | spill ld B | static smem B |")
print("|---|---|---|---|---|---|---|")
for r, d in zip(rows, dem):
    print("| %s | `%s` | %d | %d | %d | %d | %d |" % (r[0], d, r[2], r[3], r[4], r[5], r[6]))
