"""Static SASS evidence table: tensor-core / TMA / TMEM instructions per kernel of libcffm_b200.so.
usage: python tools/sass_evidence.py vss_cffm_b200/lib/libcffm_b200.so > profiles/r02_sass_evidence.md"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
cols = [("UTCHMMA", "UTCHMMA (tcgen05.mma)"), ("UTCBAR", "UTCBAR (tcgen05.commit)"), ("LDTM", "LDTM (tcgen05.ld)"), ("STTM", "STTM (tcgen05.st)"),
        ("UTMALDG", "UTMALDG (TMA load)"), ("UTMASTG", "UTMASTG (TMA store)"), ("HMMA", "HMMA (mma.sync)"), ("LDGSTS", "LDGSTS (cp.async)"),
        ("SYNCS", "SYNCS (mbarrier)"), ("FHFMA", "FHFMA (f16 x f16 + f32)"), ("MUFU", "MUFU"), ("BRA.U.ANY", "BRA.U.ANY (per-thread issue loop)")]
print("# SASS evidence (cuobjdump -sass vss_cffm_b200/lib/libcffm_b200.so, sm_100a): tensor-core / TMA / TMEM instructions per kernel\n")
print("Static instruction counts (unrolled code; loops execute them many times).  GEMM, implicit-GEMM convolution, MHA, CFM attention and the")
print("Mix-FFN tail are tcgen05 / TMEM / TMA kernels; only the fall-back MHA (N_kv > 512) uses `mma.sync`.  `BRA.U.ANY` = 0 in every producer / MMA")
print("role: each TMA / MMA issue sits under `elect.sync` in a whole-warp loop (DESIGN.md section 4, machine facts); the single one left in the")
print("fp16-only GEMM epilogue is the per-warp TMA store, whose bulk group belongs to lane 0.\n")
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---:|" * len(cols))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = dem(f.split("\n")[0].strip())
    name = re.sub(r"cffm::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((int|bool)\)", "", re.sub(r"cffm::<unnamed>::", "", name))
    name = re.sub(r"^void ", "", name).split("(")[0]
    if not re.search(r"gemm_tcgen05|mha_|cfm_attention|mixffn|patch_embed", name):
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    c = collections.Counter()
    for o in ops:
        for key, _ in cols:
            if o.startswith(key):
                c[key] += 1
    print(f"| `{name}` | " + " | ".join(str(c[k]) for k, _ in cols) + " |")
