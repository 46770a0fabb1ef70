"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel of the shipped library.
Usage: python tools/sass_table.py [path/to/libsba_attn.so] > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "sba_gan_b200", "lib", "libsba_attn.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FFMA", "BAR", "RED", "ATOM"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = name.replace("(int)", "").replace("(bool)", "").replace("(anonymous namespace)", "<unnamed>")
        name = re.sub(r"\(.*", "", name).replace("sba::<unnamed>::", "").replace("void ", "")
        order.append(name)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[name][w] += 1
        counts[name]["total"] += 1
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, SYNCS = mbarrier")
print(f"{'kernel':72s} " + " ".join(f"{w:>8s}" for w in WANT + ["total"]))
for n in order:
    c = counts[n]
    print(f"{n[:72]:72s} " + " ".join(f"{c[w]:8d}" for w in WANT + ["total"]))
