"""Attribute the stall samples of an .ncu-rep to CUDA source lines: nvdisasm line info of the
kernel's cubin (extracted from libsba_attn.so) is matched to ncu's SASS rows by position.
Usage: python tools/ncu_lines.py report.ncu-rep <cubin name, e.g. attn_tc5_bwd> <mangled-name substring> [top]"""
import csv, io, os, re, subprocess, sys, collections, tempfile
rep, cub, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "sba_gan_b200/lib/libsba_attn.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(cub)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# split per function
funcs = re.split(r"\n\s*\.section\s+\.text\.", dis)
body = [f for f in funcs if f.split("\n", 1)[0].find(ksub) >= 0]
assert body, "kernel not found in cubin"
body = body[0]
lines_of = []          # per SASS instruction: source line
cur = None
for ln in body.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blk = [b for b in src.split('"Kernel Name"')[1:] if True][0]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
print(f"sass rows: ncu {len(data)} nvdisasm {len(lines_of)}")
n = min(len(data), len(lines_of))
agg = collections.defaultdict(lambda: [0, 0])
for r, l in zip(data[:n], lines_of[:n]):
    agg[l][0] += int(r[ix['# Samples']]); agg[l][1] += int(r[ix['Instructions Executed']])
ts = sum(v[0] for v in agg.values())
srcs = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(root, "sba_gan_b200/csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:90] if l - 1 < len(srcs[f]) else ""
    print(f"{v[0] / ts:6.1%} {v[1]:>9} exec  {f}:{l}  {text}")
