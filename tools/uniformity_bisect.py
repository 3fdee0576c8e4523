#!/usr/bin/env python
"""Compile-only bisection of ptxas' uniform-datapath decision for qls_rs_pass_kernel (no GPU needed).

ptxas keeps the step loop of the register-stage kernel on the uniform datapath -- pass description through LDCU,
gate matrices as UR operands of DFMA / FFMA -- only while it can prove the loop warp-uniform, and the decision is
all-or-nothing per kernel.  When a change loses it, this tool finds which step of the
`switch` is responsible: it compiles rs_pass.cu once per step kind with every OTHER optional step compiled out
(-DQLS_RS_NO_<STEP>) and reports the FMA / FMA-with-UR counts, spills and register use of each build.

    python tools/uniformity_bisect.py                 # fp64 default instantiation, one step kind at a time
    python tools/uniformity_bisect.py --precision fp32
    python tools/uniformity_bisect.py --all           # every step enabled (what the product build contains)
    python tools/uniformity_bisect.py --extra=-DEXP=3 # pass more flags (for #if EXP experiments in the source)
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qc-quantum-linear-systems_b200", "csrc")
STEPS = ["LOADS", "TRANS", "STORES", "ACQ", "DIAGW", "DIAG", "MUX", "G1", "KR5"]  # G2 + LOAD_D/STORE_D always stay


def build(flags, precision, tmp):
    obj = os.path.join(tmp, "rs.o")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-Xptxas", "-v", "-DQLS_RS_ONLY_" + ("F64" if precision == "fp64" else "F32"),
           *flags, "-c", os.path.join(CSRC, "rs_pass.cu"), "-o", obj]
    log = subprocess.run(cmd, capture_output=True, text=True)
    if log.returncode:
        sys.exit(log.stdout + log.stderr)
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    want = "kernelId" if precision == "fp64" else "kernelIf"
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1) if ("qls_rs_pass_kernel" in m.group(1) and want in m.group(1)) else None
            if name:
                out[name] = dict(fma=0, fma_ur=0, ldc=0, ldcu=0)
            continue
        if not name:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if not m:
            continue
        op = m.group(1)
        if op in ("DFMA", "FFMA"):
            out[name]["fma"] += 1
            out[name]["fma_ur"] += bool(re.search(r"\bUR\d+\b", line))
        out[name]["ldc"] += op == "LDC"
        out[name]["ldcu"] += op == "LDCU"
    info = {}
    for m in re.finditer(r"Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s*: Used (\d+) registers", log.stdout + log.stderr):
        info[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    return out, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--extra", action="append", default=[])
    args = ap.parse_args()
    configs = [("all steps", [])] if args.all else \
        [("none (LOAD_D/STORE_D/G2 only)", [f"-DQLS_RS_NO_{s}" for s in STEPS])] + \
        [(f"+ {keep}", [f"-DQLS_RS_NO_{s}" for s in STEPS if s != keep]) for keep in STEPS]
    with tempfile.TemporaryDirectory() as tmp:
        for label, flags in configs:
            out, info = build(flags + args.extra, args.precision, tmp)
            for k, c in out.items():
                regs, st, ld = info.get(k, (0, 0, 0))
                tag = re.search(r"kernelI(\w)Li(\d+)ELi(\d+)ELi(\d+)ELi(\d+)", k)
                inst = "<%s,T=%s,KR=%s,CTAS=%s,GROUPS=%s>" % tag.groups() if tag else k
                print(f"{label:32s} {inst:34s} FMA {c['fma']:6d}  with UR {c['fma_ur']:6d}  LDC {c['ldc']:5d}  LDCU {c['ldcu']:5d}  "
                      f"regs {regs:3d}  spill {st}/{ld} B", flush=True)


if __name__ == "__main__":
    main()
