"""Per-kernel SASS opcode census of the built objects (cuobjdump -sass): proves which kernels carry tcgen05 (UTCHMMA /
UTCQMMA), TMEM loads (LDTM), TMA (UTMALDG), cp.async (LDGSTS), packed fp32 FMAs (FFMA2), MUFU etc.
usage: python scripts/sass_opcodes.py > profiles/rNN/sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "LDGSTS", "FFMA2", "FFMA", "HFMA2", "MUFU",
         "LDS", "STS", "LDG", "STG", "RED", "ATOM", "ELECT", "ACQBULK", "BAR"]
print("# per-kernel SASS opcode counts (static instruction counts, cuobjdump -sass, sm_100a), objects under crfp_b200/_obj")
print("# %-64s %s" % ("kernel", " ".join(f"{w:>8s}" for w in WATCH)))
for obj in sorted(glob.glob(os.path.join(ROOT, "crfp_b200", "_obj", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn).replace("void ", "").replace("crfp::", "")
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            op = m.group(1)
            for w in WATCH:
                if op == w or (w in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "LDGSTS", "MUFU", "SYNCS", "UTCBAR") and op.startswith(w)):
                    counts[fn][w] += 1
                    break
    print(f"## {os.path.basename(obj)}")
    for fn, c in counts.items():
        if sum(c.values()) == 0:
            continue
        print("  %-64s %s" % (fn[:64], " ".join(f"{c.get(w, 0):8d}" for w in WATCH)))
