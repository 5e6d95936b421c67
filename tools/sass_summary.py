"""Counts the SASS mnemonics that prove (or disprove) a Blackwell-native path, per object file of the shipped library:
    python tools/sass_summary.py > profiles/r2_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, HMMA = legacy mma.sync,
MULTIMEM = NVLS multimem.ld_reduce / st, STG.E.EF.128 = 128-bit streaming stores (B200_PROFILING.md)."""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                               ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("MULTIMEM", r"\bMULTIMEM|\bLDGMC|\bSTGMC|\.MC\b"),
                               ("STG.E.EF.128", r"STG\.E\.EF\.128"), ("LDS.128", r"\bLDS\.128"), ("SYNCS", r"\bSYNCS")])
print("object".ljust(16) + "".join(k.rjust(14) for k in PAT) + "   kernels")
for obj in sorted(glob.glob(os.path.join(ROOT, "maexp_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels = len(re.findall(r"^\s*Function :", sass, re.M))
    print(os.path.basename(obj).ljust(16) + "".join(str(len(re.findall(p, sass))).rjust(14) for p in PAT.values()) + f"   {kernels}")
