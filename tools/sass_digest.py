"""cuobjdump -sass digest of the shipped libvmorph.so: per kernel, how often the instructions that identify the sm_100a features
in use appear.  Writes profiles/r2_sass_digest.md.    python tools/sass_digest.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "videomorphing_b200", "libvmorph.so")
KEYS = ["UTMALDG", "SYNCS", "UCGABAR", "FFMA2", "FADD2", "FMUL2", "CCTL", "MUFU", "SHFL", "ATOM", "RED", "MEMBAR", "ERRBAR", "BAR", "LDG", "LDS", "STS", "FFMA", "FADD", "FMUL", "IMAD"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)[^\n]*SHARED:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op == k or (k in ("ATOM", "RED", "BAR", "UCGABAR", "SYNCS", "UTMALDG", "CCTL") and op.startswith(k)):
                    kernels[cur][k] += 1
    dem = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS digest of videomorphing_b200/libvmorph.so (sm_100a) — `python tools/sass_digest.py`", "",
           "Static instruction counts per kernel (`cuobjdump -sass`); registers / static shared memory from `cuobjdump -res-usage`.",
           "`UTMALDG` = TMA tensor load (`cp.async.bulk.tensor`), `SYNCS` = mbarrier, `UCGABAR` = hardware cluster barrier, `FFMA2 / FADD2 / FMUL2` = packed",
           "fp32x2 arithmetic, `CCTL` = L1 prefetch (`prefetch.global.L1`), `MUFU` = rcp / rsqrt of the exact division / square-root sequences.", "",
           "| kernel | regs | smem B | instr | " + " | ".join(KEYS) + " |", "|---|---|---|---|" + "---|" * len(KEYS)]
    for (k, c), d in zip(kernels.items(), dem):
        name = re.sub(r"\(.*", "", d).replace("void ", "")
        r = regs.get(k, ("", ""))
        out.append(f"| `{name}` | {r[0]} | {r[1]} | {c['_total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
    open(os.path.join(ROOT, "profiles", "r2_sass_digest.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
