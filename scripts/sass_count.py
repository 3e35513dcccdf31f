"""Static SASS statistics of one object file: per kernel, total instructions and the length of every loop (backward branch),
with the mix of a few instruction classes.  Usage: python scripts/sass_count.py build/online_loop_ws.o [name filter]"""
import re
import subprocess
import sys
from collections import Counter

obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name, body = f.split("\n", 1)
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    if flt not in dem:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", body)
    addrs = [int(a, 16) for a, _ in ins]
    ops = [o for _, o in ins]
    print("== %s: %d instructions" % (dem[:110], len(ins)))
    for i, (a, o) in enumerate(zip(addrs, ops)):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", o)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            j = addrs.index(t) if t in addrs else None
            if j is None:
                continue
            seg = ops[j:i + 1]
            c = Counter()
            for s in seg:
                s = re.sub(r"^@!?U?P\d+\s+", "", s)
                c[s.split()[0].split(".")[0]] += 1
            top = ", ".join("%s %d" % kv for kv in c.most_common(14))
            print("   loop %#x..%#x: %d instr  [%s]" % (t, a, len(seg), top))
