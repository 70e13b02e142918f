"""Static SASS mnemonic counts per kernel of the built library (no GPU needed):

    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "neural_audio_tokenizer_b200", "libnat_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "STG.E.ENL2.256", "STG.E.256", "LDG.E.ENL2.256",
        "LDG.E.256", "MUFU.LG2", "MUFU.SQRT", "DFMA", "F2F.F64.F32", "FFMA2", "STS.64", "LDS.64", "LDS.128", "STS.128", "STL", "LDL"]


def main():
    import bench
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print(f"# cuobjdump -sass neural_audio_tokenizer_b200/libnat_b200.so (build {bench.build_id()} of csrc/, sm_100a): instruction and")
    print("# mnemonic counts per kernel (static counts, not executed counts). UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM = tcgen05.ld,")
    print("# UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, *.ENL2.256 = the 256-bit global accesses as")
    print("# cuobjdump prints them, STL / LDL = spill traffic.")
    print("# Produced by profiles/sass_summary.py.")
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        short = re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", f, flags=re.M))
        counts = [(k, len(re.findall(r"\b" + re.escape(k) + r"\b", f))) for k in KEYS]
        print(f"{short}: {n} instructions; " + ", ".join(f"{k} {v}" for k, v in counts if v))


if __name__ == "__main__":
    main()
