"""Opcode histogram of every kernel in libtvit_b200.so (run in the build container: cuobjdump needs no GPU).

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt

Evidence for "what proves a Blackwell-native kernel" (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, packed fp32 -> FFMA2/FMUL2/FADD2, legacy mma.sync -> HMMA (must be 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neural_vit_b200", "libtvit_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FFMA2", "FMUL2", "FADD2",
       "MUFU", "REDG", "ATOMG", "ATOMS", "HMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(names)} kernels; columns = total instructions, then the key opcodes")
    print("# " + " ".join(KEY))
    for n, d in zip(names, dem):
        c = kernels[n]
        total.update(c)
        short = re.sub(r"\(.*", "", d).replace("tvit::", "")
        keys = " ".join(f"{k}={c[k]}" for k in KEY if c[k])
        print(f"{short:70s} n={sum(c.values()):6d}  {keys}")
    print("\n# whole library")
    print(" ".join(f"{k}={total[k]}" for k in KEY))
    top = ", ".join(f"{k}:{v}" for k, v in total.most_common(25))
    print("# most frequent opcodes: " + top)


if __name__ == "__main__":
    sys.exit(main())
