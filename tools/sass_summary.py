"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass), for profiles/.
    python tools/sass_summary.py [path/to/libsafe_b200.so]
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA engine), UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UBLKCP[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTCBAR[.\w]*|"
                 r"SYNCS[.\w]*|HMMA[.\w]*|IMMA[.\w]*|LDGSTS[.\w]*|PRMT|UTCATOMSWS[.\w]*|UTCCP[.\w]*)\b")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "safepy_b200", "libsafe_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(ln.split("/*")[1] if ln.count("/*") >= 2 else ln)
        if m:
            per[cur][m.group(1).split(".")[0] if m.group(1).startswith("SYNCS") else m.group(1)] += 1
    total = collections.Counter()
    print("library: %s" % os.path.relpath(lib, ROOT))
    print("built for: %s" % ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
    for name, cnt in per.items():
        keys = [k for k in cnt if not k.startswith(("SYNCS", "PRMT"))]
        if not keys:
            continue
        short = re.sub(r"\(.*", "", name)
        print("%-58s %s" % (short[:58], "  ".join("%s x%d" % (k, cnt[k]) for k in sorted(cnt))))
        total.update(cnt)
    print("TOTAL over tcgen05 / TMA kernels: " + "  ".join("%s x%d" % (k, total[k]) for k in sorted(total)))


if __name__ == "__main__":
    main()
