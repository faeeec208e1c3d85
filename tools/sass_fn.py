"""Extracts one function's SASS from a library (cuobjdump -sass) — instructions only, one per line.
usage: sass_fn.py lib.so k_stage_pass [--loop]   --loop: print the node round (the block around the 7 node-record loads)"""
import re, subprocess, sys
lib, name = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, out = None, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and name in fn:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
print("# %s: %d instructions" % (name, len(out)))
if "--loop" in sys.argv:
    # the node round: 7 consecutive-ish LDG.E.128 with a cache-hint descriptor
    idx = [i for i, (a, s) in enumerate(out) if "LDG.E.128" in s]
    runs = []
    for i in idx:
        if runs and i - runs[-1][-1] <= 3:
            runs[-1].append(i)
        else:
            runs.append([i])
    for r in runs:
        if len(r) >= 6:
            print("# run of %d 128-bit loads at instruction %d (addr %#x)" % (len(r), r[0], out[r[0]][0]))
else:
    for a, s in out:
        print("%06x  %s" % (a, s))
