"""Copies the round-2 evidence from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, pytest log, ncu launch list +
per-kernel share summary, per-launch `ncu --set full` one-liners, the dram-bytes table bench.py reads, and a SASS opcode count
per kernel of the in-tree library (tcgen05 / TMEM / TMA evidence)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def copy(name, dst=None):
    src = os.path.join(O, name)
    if os.path.exists(src) and os.path.getsize(src) > 0:
        shutil.copy(src, os.path.join(P, dst or name))
        return True
    print("missing", name)
    return False


for n in ("r02_bench.json", "r02_bench_reference_arm.json", "r02_bench_cfg5_rfb320_allpriors.json", "r02_bench_cfg5_rfb640_b64_p1.json",
          "r02_bench_cfg5_rfb640_b64_p5.json", "r02_bench_cfg5_rfb640_b64_p50.json", "r02_pytest_gpu.log", "r02_ncu_launches_bench.csv",
          "r02_bench_2gpu.json", "r02_jpeg_profile.txt", "r02_ncu_jpeg_per_launch.txt", "r02_ncu_jpeg_enc_per_launch.txt", "r02_bench_4gpu.json", "r02_bench_8gpu.json"):
    copy(n)

# per-kernel share of the step from the launch list (cold-cache, serialised: shares, not absolutes)
p = os.path.join(O, "r02_ncu_launches_bench.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(l for l in open(p) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"<.*", "", r[ik].replace("void ", "").replace("uf::", "").split("(")[0])
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values()) or 1.0
    with open(os.path.join(P, "r02_ncu_launch_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 2 --warmup 3` (cold cache, serialised:\n"
                "# the SHARE per kernel is what compares with bench.py's event timing, not the absolute)\n")
        for k, v in tot.most_common():
            f.write("%-34s launches %5d  time %10.1f us  share %5.1f %%\n" % (k, cnt[k], v / 1e3, 100 * v / s))

for n in ("r02_ncu_full_per_launch.txt", "r02_ncu_dram_bytes.csv", "r02_ncu_tensor_pipe.txt"):  # summarised on the GPU box (tools/r02_evidence.sh)
    copy(n)

# SASS evidence from the in-tree library
lib = os.path.join(ROOT, "infercam_onnx_b200", "libultraface_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cur.replace("void ", "").replace("uf::", "")
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for op in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "FFMA", "SYNCS", "REDUX", "VOTE"):
            if re.search(r"\b%s\b" % op, line) or (" %s." % op) in line:
                counts[cur][op] += 1
with open(os.path.join(P, "r02_sass_opcode_counts.txt"), "w") as f:
    f.write("# cuobjdump -sass infercam_onnx_b200/libultraface_b200.so (sm_100a): static instruction counts per kernel\n"
            "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n")
    ops = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "REDUX", "VOTE", "FFMA")
    f.write("%-64s %s\n" % ("kernel", " ".join("%8s" % o for o in ops)))
    tot = collections.Counter()
    for k, c in counts.items():
        f.write("%-64s %s\n" % (k[:64], " ".join("%8d" % c[o] for o in ops)))
        tot.update(c)
    f.write("%-64s %s\n" % ("TOTAL", " ".join("%8d" % tot[o] for o in ops)))
print("profiles/ updated:", sorted(n for n in os.listdir(P) if n.startswith("r02")))
