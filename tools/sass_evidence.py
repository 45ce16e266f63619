"""Per-kernel SASS evidence for profiles/: disassembles the in-tree librd_b200.so (cuobjdump -sass) and, for every convolution kernel,
counts the Blackwell-only mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores,
UTCBAR = tcgen05.commit, REDG...F32x4 = vector reductions) and prints the first occurrences with their addresses.
    python tools/sass_evidence.py > profiles/r02_sass_conv_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "representation-disentanglement_b200", "librd_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|SYNCS|REDG\.E\.ADD\.F32x4|UTCATOMSWS")
want = re.compile(r"k_conv_halo|k_conv_tma|k_wgrad_halo|k_wgrad_tma|k_conv_tc|k_wgrad_tc")
print("cuobjdump -sass %s  (sm_100a)" % os.path.relpath(so, ROOT))
print("arch of the embedded cubins:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", txt)))))
cur, body = None, []
funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
tot = collections.Counter()
for name, lines in funcs.items():
    if not want.search(name):
        continue
    try:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        dem = name
    cnt = collections.Counter()
    first = {}
    n_instr = 0
    for l in lines:
        mm = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if not mm:
            continue
        n_instr += 1
        ins = mm.group(2).strip()
        k = pat.search(ins)
        if k:
            key = k.group(0)
            cnt[key] += 1
            first.setdefault(key, []).append("    /*%s*/ %s" % (mm.group(1), ins))
    tot.update(cnt)
    print("\n== %s\n   %d SASS instructions; %s" % (re.sub(r"\(anonymous namespace\)::", "", dem)[:150], n_instr,
                                                  ", ".join("%s x%d" % kv for kv in sorted(cnt.items()))))
    for key in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "REDG.E.ADD.F32x4"):
        for ex in first.get(key, [])[:2]:
            print(ex)
print("\nlibrary totals over these kernels:", ", ".join("%s x%d" % kv for kv in sorted(tot.items())))
