"""Static evidence that the kernels use the Blackwell paths: counts of tensor-core / TMA / TMEM / mbarrier SASS
mnemonics per kernel of the built library (CPU only: cuobjdump on the cross-compiled .so).

  python tools/sass_mnemonics.py [lib.so] [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|UTCBAR|UTCATOMSWS|SYNCS|UCGABAR|ELECT|USETMAXREG)")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mofanerf_b200", "lib", "libmofa_b200.so")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r01_sass_mnemonics.txt")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if cur is None or not m:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", m.group(1)).split()[0]
        if PAT.match(op):
            cnt[cur]["SYNCS.*" if op.startswith("SYNCS") else op] += 1
    out = ["# cuobjdump -sass " + os.path.relpath(lib, ROOT) + ": static counts per kernel",
           "# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2); UTMALDG / UTMASTG = cp.async.bulk.tensor load / store;",
           "# UTMAPF = cp.async.bulk.prefetch.tensor; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; UTCATOMSWS = tcgen05.alloc;",
           "# SYNCS.* = mbarrier ops; USETMAXREG = setmaxnreg; UCGABAR = cluster barrier; ELECT = elect.sync"]
    for k, c in cnt.items():
        if not any(x.startswith(("UTC", "UTMA", "LDTM")) for x in c):
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        out.append(re.sub(r"\(.*", "", name))
        out.append("    " + ", ".join(f"{a} x{b}" for a, b in sorted(c.items())))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
