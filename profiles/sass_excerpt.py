"""SASS of the innermost loops (backward branches whose body holds packed FMA work) of one kernel of the shipped library:
    python profiles/sass_excerpt.py <mangled-name-substring> [min FFMA2 in the loop] > profiles/r02/sass_<kernel>.txt"""
import re
import subprocess
import sys

so = "point-dae_b200/lib/libpointdae_b200.so"
pat = sys.argv[1]
min_ffma = int(sys.argv[2]) if len(sys.argv) > 2 else 8
names = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, blocks = None, {}
for line in names.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        blocks[cur] = []
    elif cur:
        blocks[cur].append(line)
hits = [n for n in blocks if pat in n]
for name in hits[:1]:
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    ins = []
    for l in blocks[name]:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("kernel:", demangled)
    print("instructions:", len(ins))
    census = {}
    for _, t in ins:
        op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
        census[op] = census.get(op, 0) + 1
    print("TMA / mbarrier evidence:", " ".join("%s:%d" % (o, sum(1 for _, t in ins if o in t)) for o in ("UBLKCP", "SYNCS", "UTMALDG", "UTCHMMA", "UTCMMA", "LDTM")))
    print("whole-kernel census:", " ".join("%s:%d" % x for x in sorted(census.items(), key=lambda x: -x[1])[:24]))
    for pc, t in ins:
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2 and int(m2.group(1), 16) < pc:
                tgt = int(m2.group(1), 16)
                body = [x for x in ins if tgt <= x[0] <= pc]
                ops = {}
                for _, b in body:
                    op = (b.split()[1] if b.startswith("@") else b.split()[0]).split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                if ops.get("FFMA2", 0) >= min_ffma and len(body) < 700:
                    print("\n==== loop 0x%x .. 0x%x: %d instructions: %s" % (
                        tgt, pc, len(body), " ".join("%s:%d" % x for x in sorted(ops.items(), key=lambda x: -x[1]))))
                    for p, b in body:
                        print("  /*%04x*/ %s" % (p, b))
