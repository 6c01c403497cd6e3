"""Development aid: for every innermost loop of a kernel that contains a MUFU.RSQ64H (= one normal_pair), print its
instruction count.   python tools/sass_loops.py <mangled-name-substring>"""
import re, subprocess, sys
so = "multigridmc_b200/csrc/libmgmc_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name, body = b.split("\n", 1)
    if sys.argv[1] not in name: continue
    lines = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m: lines.append((int(m.group(1), 16), m.group(2).strip()))
    addr2idx = {a: i for i, (a, _) in enumerate(lines)}
    loops = []
    for i, (a, ins) in enumerate(lines):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m:
            t = int(m.group(1), 16)
            if t <= a and t in addr2idx: loops.append((addr2idx[t], i))
    print(name, len(lines), "instructions")
    for (s, e) in loops:
        body_ins = [x[1] for x in lines[s:e + 1]]
        n = sum(1 for x in body_ins if "MUFU.RSQ64H" in x)
        if n and not any(s < s2 and e2 < e for (s2, e2) in loops if (s2, e2) != (s, e)):
            ops = {}
            for x in body_ins:
                p = x.split()
                if p[0].startswith("@"): p = p[1:]
                ops[p[0].split(".")[0]] = ops.get(p[0].split(".")[0], 0) + 1
            print(f"  loop {s}-{e}: {e - s + 1} instr, {n} rsq;", dict(sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
