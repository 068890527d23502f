"""SASS opcode census of libinference_engine.so (runs on the CPU box: `cuobjdump -sass`): per kernel, how many tcgen05 MMAs
(UTC*MMA), TMA loads/stores (UTMALDG / UTMASTG), tensor-memory loads/stores (LDTM / STTM), tcgen05 commits (UTCBAR), legacy
tensor-core instructions (HMMA) and plain FP32 FMAs it contains.  usage: python tools/sass_census.py > profiles/<tag>_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpu-ai-inference-server_b200", "lib", "libinference_engine.so")
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
                               ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("FFMA", r"\bFFMA"),
                               ("F2FP", r"\bF2FP")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    kinds = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            kinds[cur] = set()
            continue
        if cur is None:
            continue
        for k, pat in PAT.items():
            mm = re.search(pat + r"[.\w]*", line)
            if mm:
                counts[cur][k] += 1
                if k == "UTC*MMA":
                    kinds[cur].add(mm.group(0).split(".")[0])
    names = list(counts)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dm):
        demangle[n] = d
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# " + " ".join(f"{k:>8s}" for k in PAT) + "  kernel")
    for n in names:
        c = counts[n]
        short = re.sub(r"\(anonymous namespace\)::|b200::kernels::|<unnamed>::", "", demangle.get(n, n))
        short = re.sub(r"\((int|bool|unsigned int)\)", "", short)
        short = re.sub(r"\(.*", "", short)
        print("  " + " ".join(f"{c[k]:8d}" for k in PAT) + f"  {short[:110]}" + (f"   [{','.join(sorted(kinds[n]))}]" if kinds[n] else ""))
    tc = [n for n in names if counts[n]["UTC*MMA"]]
    print(f"# {len(names)} kernels, {len(tc)} of them issue tcgen05.mma; legacy HMMA instructions in the library: {sum(counts[n]['HMMA'] for n in names)}")


if __name__ == "__main__":
    sys.exit(main())
