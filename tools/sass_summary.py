#!/usr/bin/env python
"""Static evidence for profiles/: per hot kernel of libhx_b200.so the register count / spills (ptxas -v) and
the SASS instruction mix (cuobjdump -sass): FP64 tensor-core DMMA, system-scope flag loads/stores of the
peer-memory kernels, streaming (no-allocate) global loads.  Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOT = ("tail_kernel", "sell_kernel", "spmv_csr_kernel", "jacobi_kernel", "multi_dot_kernel", "multi_axpy_kernel", "scale_copy_kernel",
       "basis_rotate", "halo_exchange_kernel", "allreduce_kernel", "color_round_kernel", "assemble_AC_kernel",
       "dense_gemv_kernel", "spgemm_numeric_kernel")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    from helmholtz_x_b200 import build
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    res = subprocess.run([nvcc] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-o", "/tmp/hx_sass_probe.so"] + build.SRC,
                         capture_output=True, text=True)
    regs = {}
    cur = None
    for ln in res.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"Used (\d+) registers.*?(\d+) bytes smem|Used (\d+) registers", ln)
        if m and cur:
            regs[cur] = ln.split(":", 1)[1].strip()
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and cur:
            regs[cur + "#spill"] = f"spill stores {m.group(2)} B, loads {m.group(3)} B"
    sass = subprocess.run(["cuobjdump", "-sass", build.OUT], capture_output=True, text=True).stdout
    mix = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            mix[cur][m.group(1)] += 1
    names = demangle(list(mix))
    print("# libhx_b200.so: static evidence per hot kernel (sm_100a, nvcc 12.9, -O3 -lineinfo)\n")
    print("`python tools/sass_summary.py` -- ptxas -v and `cuobjdump -sass`, no GPU needed.  DMMA = FP64 tensor-core "
          "mma.sync.m8n8k4; `.STRONG.SYS` = system-scope acquire/release accesses (flags in peer memory); "
          "`LDG.E...CONSTANT`/`.NA` = streaming read-only loads.\n")
    print("| kernel | registers / smem | spills | instructions | DMMA | DFMA | LDG | STG | system-scope LD/ST | notes |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for mangled, c in mix.items():
        nm = names[mangled]
        if not any(h in nm for h in HOT):
            continue
        total = sum(c.values())
        dmma = sum(v for k, v in c.items() if k.startswith("DMMA"))
        dfma = sum(v for k, v in c.items() if k.startswith("DFMA") or k.startswith("DMUL") or k.startswith("DADD"))
        ffma = sum(v for k, v in c.items() if k.startswith("FFMA") or k.startswith("FMUL") or k.startswith("FADD"))
        ldg = sum(v for k, v in c.items() if k.startswith("LDG"))
        stg = sum(v for k, v in c.items() if k.startswith("STG"))
        sysacc = sum(v for k, v in c.items() if ".SYS" in k and (k.startswith("LD") or k.startswith("ST")))
        na = sum(v for k, v in c.items() if k.startswith("LDG") and ("CONSTANT" in k or ".NA" in k))
        note = []
        if ffma and not dfma:
            note.append(f"complex64 path ({ffma} FP32 ops)")
        if na:
            note.append(f"{na} read-only/streaming LDG")
        short = re.sub(r"\(.*", "", nm).replace("void hx::", "").replace("hx::", "")
        print(f"| `{short}` | {regs.get(mangled, '?')} | {regs.get(mangled + '#spill', '-')} | {total} | {dmma} | {dfma} | {ldg} | {stg} | {sysacc} | {'; '.join(note)} |")


if __name__ == "__main__":
    main()
