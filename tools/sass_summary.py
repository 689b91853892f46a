"""Per-kernel SASS mnemonic counts of the built libay2.so (cuobjdump -sass): the evidence that the conv kernels are
tcgen05 / TMEM / TMA code (B200_PROFILING.md "What proves a Blackwell-native kernel"). Writes profiles/<tag>_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ayolov2_b200", "libay2.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA.", "LDGSTS", "MUFU", "RED.", "ATOM"]


def main(tag: str) -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip() or n  # noqa: E731
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w.rstrip(".")) and (not w.endswith(".") or op.startswith(w) or op == w.rstrip(".")):
                    if w == "HMMA." and op.startswith("UTCHMMA"):
                        continue
                    per[cur][w.rstrip(".")] += 1
    out = [f"# SASS mnemonic counts per kernel of ayolov2_b200/libay2.so (cuobjdump -sass; tag {tag}).",
           "# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA loads / stores -> UTMALDG / UTMASTG, tcgen05.commit -> UTCBAR;",
           "# HMMA would be the legacy mma.sync path (none expected).", ""]
    cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "MUFU", "RED", "ATOM", "_total"]
    out.append(f"{'kernel':90s} " + " ".join(f"{c:>8s}" for c in cols))
    tot = collections.Counter()
    for name, cnt in per.items():
        pretty = demangle(name)
        i = pretty.rfind(">(")
        pretty = pretty[:i + 1] if i >= 0 else pretty.split("(")[0]
        pretty = pretty.replace("(bool)", "").replace("(int)", "").replace("void ", "")[:90]
        out.append(f"{pretty:90s} " + " ".join(f"{cnt.get(c, 0):8d}" for c in cols))
        tot.update(cnt)
    out.append(f"{'TOTAL':90s} " + " ".join(f"{tot.get(c, 0):8d}" for c in cols))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, "kernels:", len(per), "UTCHMMA", tot["UTCHMMA"], "HMMA", tot["HMMA"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
