"""Per-kernel counts of the Blackwell-native SASS instructions in libgrasp_b200.so (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, UTCBAR = tcgen05.commit."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "grasp_b200/libgrasp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UTCBAR": r"\bUTCBAR", "HMMA (legacy)": r"\bHMMA", "SYNCS": r"\bSYNCS", "DFMA": r"\bDFMA"}
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {lib}: SASS instruction counts per kernel (kernels without any of them omitted)")
cols = list(pat)
print("kernel".ljust(72) + "".join(c.rjust(15) for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c[x] for x in cols if x not in ("SYNCS", "DFMA")) == 0:
        continue
    tot.update(c)
    print(k[-70:].ljust(72) + "".join(str(c[x]).rjust(15) for x in cols))
print("TOTAL".ljust(72) + "".join(str(tot[x]).rjust(15) for x in cols))
