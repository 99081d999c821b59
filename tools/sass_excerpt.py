"""SASS evidence per kernel of libhct_b200.so: counts of the Blackwell-native mnemonics (cuobjdump -sass), written to
profiles/.  Runs without a GPU.   python tools/sass_excerpt.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "headct_foundation_b200", "lib", "libhct_b200.so")
PATTERNS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCBAR", "REDG", "MUFU.EX2", "MUFU.TANH",
            "HMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        for p in PATTERNS:
            if p in line:
                # keep the variant suffix (.2CTA, .MULTICAST, .x32 ...) of the first token that contains the pattern
                tok = next((t for t in re.split(r"[\s,;]+", line) if p in t), p)
                cur[tok] += 1
    out = ["SASS mnemonics per kernel of headct_foundation_b200/lib/libhct_b200.so (cuobjdump -sass; tools/sass_excerpt.py)",
           "tcgen05.mma -> UTCHMMA, TMA load / store -> UTMALDG / UTMASTG, bulk copy -> UBLKCP, tcgen05.ld / st -> LDTM / STTM", ""]
    for name, cnt in kernels.items():
        if not cnt:
            continue
        out.append(name)
        out.append("    " + "  ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
    txt = "\n".join(out) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt")
    with open(dst, "w") as f:
        f.write(txt)
    print(txt[:3000])


if __name__ == "__main__":
    main()
