#!/usr/bin/env python
"""Opcode evidence for the shipped library: per kernel of libwdg.so, how many tcgen05 / TMA / TMEM instructions its
sm_100a SASS holds (`cuobjdump -sass`, mnemonics per /opt/skills/guides/B200_PROFILING.md).  Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wind_downscaling_gan_b200", "csrc", "libwdg.so")
PATTERNS = [("tcgen05.mma", r"\bUTC[A-Z]*MMA\b"), ("of which cta_group::2", r"\bUTC[A-Z]*MMA\.2CTA\b"), ("TMA load", r"\bUTMALDG\b"),
            ("tcgen05.ld", r"\bLDTM\b"),
            ("tcgen05.commit/barrier", r"\bUTCBAR\b"), ("TMEM alloc", r"\bUTCATOMSWS\b"), ("cluster barrier", r"\bUCGABAR_ARV\b"),
            ("mbarrier", r"\bSYNCS\b"),
            ("legacy HMMA", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    assert len(blocks) == len(names)
    rows = []
    total = collections.Counter()
    for name, body in zip(names, blocks):
        c = {k: len(re.findall(p, body)) for k, p in PATTERNS}
        total.update(c)
        if c["tcgen05.mma"] or c["TMA load"]:
            short = name.replace("(anonymous namespace)::", "").replace("wdg::", "").replace("void ", "")
            short = short.split(">(")[0] + ">" if ">(" in short else short.split("(")[0]
            rows.append((short, c))
    print("# SASS opcode summary of libwdg.so (sm_100a), round 2\n")
    print(f"`cuobjdump -sass {os.path.relpath(LIB, ROOT)}`: {len(names)} kernels, {len(rows)} of them issue tcgen05 / TMA instructions.")
    print("Template arguments of the inference kernels end in the operand precision: `0` = bf16 (`kind::f16`), `1` = tf32 "
          "(`kind::tf32`); both kinds assemble to `UTCHMMA` (the kind lives in the instruction descriptor).  `conv_pair_kernel` and "
          "`halo_conv_kernel<..., true>` are the CTA-pair kernels: `UTCHMMA.2CTA`, `UTMALDG.2CTA`, `UTCBAR.2CTA.MULTICAST`.\n")
    print("| whole library | " + " | ".join(k for k, _ in PATTERNS) + " |")
    print("|---|" + "---|" * len(PATTERNS))
    print("| all kernels | " + " | ".join(str(total[k]) for k, _ in PATTERNS) + " |\n")
    print("| kernel | " + " | ".join(k for k, _ in PATTERNS[:7]) + " |")
    print("|---|" + "---|" * 7)
    for short, c in sorted(rows, key=lambda r: r[0]):
        print(f"| `{short}` | " + " | ".join(str(c[k]) for k, _ in PATTERNS[:7]) + " |")


if __name__ == "__main__":
    sys.exit(main())
