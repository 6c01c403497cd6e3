"""Development aid: dump one kernel's SASS from the built library and print an opcode histogram of a line range.
    python tools/sass_loop.py <mangled-name-substring> [first_line last_line]
"""
import collections
import re
import subprocess
import sys

so = "multigridmc_b200/csrc/libmgmc_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
sel = [b for b in blocks[1:] if sys.argv[1] in b.split("\n", 1)[0]]
assert sel, "no such kernel"
b = sel[0]
name, body = b.split("\n", 1)
lines = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        lines.append((m.group(1), m.group(2).strip()))
print(name, len(lines), "instructions")
if len(sys.argv) >= 4:
    a, z = int(sys.argv[2]), int(sys.argv[3])
    h = collections.Counter()
    for addr, ins in lines[a:z + 1]:
        p = ins.split()
        if p[0].startswith("@"):
            p = p[1:]
        h[p[0].split(".")[0]] += 1
    print(z - a + 1, "instructions in range:", dict(h.most_common()))
else:
    with open("/tmp/sass_sel.txt", "w") as f:
        for k, (addr, ins) in enumerate(lines):
            f.write(f"{k:5d} {addr} {ins}\n")
    print("written /tmp/sass_sel.txt")
