"""Digest a gpurun_out/<tag>_* profile pack (tools/profile_all.sh) into profiles/: launch shares, hot-kernel summary.
usage: python tools/profile_digest.py <tag> <out-prefix>      e.g.  r1v10 r1_v10"""
import collections, csv, io, os, shutil, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = lambda f: os.path.join(root, "gpurun_out", f)
p = lambda f: os.path.join(root, "profiles", f)
rows = [r for r in csv.reader(open(g(f"{tag}_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ik].split("(")[0]
    v = float(r[iv].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[iu], 1.0)
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
with open(p(f"{out}_launch_shares.csv"), "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, v in agg.most_common():
        f.write(f"\"{k}\",{cnt[k]},{v:.3f},{v / tot:.4f}\n")
shutil.copy(g(f"{tag}_launches.csv"), p(f"{out}_launches_ncu.csv"))
shutil.copy(g(f"{tag}_plain.json"), p(f"{out}_bench_for_launch_list.json"))
s = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), g(f"{tag}_hot.ncu-rep")], capture_output=True, text=True).stdout
open(p(f"{out}_hot_kernels_summary.txt"), "w").write(s)
d = subprocess.run(["ncu", "-i", g(f"{tag}_hot.ncu-rep"), "--page", "details"], capture_output=True, text=True).stdout
open(p(f"{out}_hot_kernels_details.txt"), "w").write(d)
print(open(p(f"{out}_launch_shares.csv")).read().splitlines()[:4])
