"""Per-kernel SASS mnemonic counts of libedgedict_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / UTCQMMA),
TMA (UTMALDG / UTMASTG), bulk DSMEM copies (UBLKCP), TMEM loads (LDTM), legacy tensor-core MMAs (HMMA), cp.async (LDGSTS),
cluster barriers (UCGABAR), mbarrier waits (SYNCS) and cta_group::2 forms (any mnemonic carrying .2CTA).  Writes profiles/r2/sass_summary.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "edgedict_b200", "libedgedict_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "LDSM", "UCGABAR", "SYNCS", "MEMBAR", "ATOMG", "RED", "2CTA"]
rows, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)[:70]
        rows.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                rows[cur][k] += 1
        if ".2CTA" in op:
            rows[cur]["2CTA"] += 1
        rows[cur]["_total"] += 1
os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r2", "sass_summary.txt"), "w") as fh:
    fh.write("SASS mnemonic counts per kernel of edgedict_b200/libedgedict_b200.so (cuobjdump -sass, sm_100a); scripts/sass_summary.py\n")
    fh.write("%-72s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in KEYS) + "\n")
    agg = collections.OrderedDict()
    for name, c in rows.items():
        base = re.sub(r"<.*", "", name)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(c)
    for base, (n, c) in agg.items():
        label = "%s  [%d instantiation%s]" % (base, n, "" if n == 1 else "s")
        fh.write("%-72s %7d " % (label[:72], c["_total"]) + " ".join("%8d" % c[k] for k in KEYS) + "\n")
print(open(os.path.join(ROOT, "profiles", "r2", "sass_summary.txt")).read())
