"""tools/sass_evidence.py -- per-kernel SASS facts of the built library: instruction counts and the mnemonics that show which
hardware paths a kernel uses (TMA + mbarrier, streaming loads / stores, FP64 fallback), plus ptxas's registers / spills.
usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdf_tools_b200", "lib", "libsdf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# SASS evidence for libsdf_b200.so (sm_100a), generated with tools/sass_evidence.py (`cuobjdump -sass`, `nvcc -Xptxas -v`)")
print("# per kernel: instruction count, selected mnemonics; TMA + mbarrier forms for the banded kernel (B200_PROFILING.md),")
print("# evict-first (.EF) loads / stores for the line-per-thread kernel\n")
cur, ops = None, None
KEYS = ["UTMALDG", "SYNCS", "ATOMS", "ATOMG", "REDG", "BAR", "LDS", "STS", "LDG", "STG", "LDL", "STL", "DFMA", "DMUL", "MUFU", "FLO", "SHFL", "WARPSYNC", "CREDUX", "IMAD.WIDE"]
def flush():
    if cur is None: return
    n = sum(ops.values())
    sel = " ".join(f"{k}={sum(v for o, v in ops.items() if o.startswith(k))}" for k in KEYS if any(o.startswith(k) for o in ops))
    print(f"{demangle(cur)[:170]}\n    instructions={n}  {sel}")
    forms = sorted(o for o in ops if o.startswith(("UTMALDG", "SYNCS")) or ".EF" in o)
    print(f"    notable forms: {forms}")
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush(); cur, ops = m.group(1), collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur: ops[m.group(1)] += 1
flush()
print("\n# ptxas -v (registers, spills, shared memory) of the envelope and binary kernels")
csrc = os.path.join(ROOT, "sdf_tools_b200", "csrc")
for f in ("edt_kernels.cu", "line_kernels.cu"):
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v", "-c", os.path.join(csrc, f), "-o", "/dev/null"],
                       capture_output=True, text=True).stderr.splitlines()
    for i, l in enumerate(r):
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            name = demangle(m.group(1))
            if not any(k in name for k in ("band_strided", "binary_z", "line_hull")): continue
            body = " ".join(x.strip() for x in r[i + 1:i + 4] if "ptxas info" in x or "bytes stack" in x)
            body = re.sub(r"ptxas info\s*:\s*", "", body)
            print(f"{name[:120]}\n    {body}")
