#!/usr/bin/env python
"""Static evidence for every kernel of the library, no GPU needed: registers / shared memory / spills from
``ptxas -v`` and the memory-instruction mix from ``cuobjdump -sass`` (widths of global loads / stores, TMA bulk copies,
mbarrier ops, atomics).  Written to profiles/ as a plain table.

    python tools/static_report.py > profiles/rN_static_sass.txt
"""
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ood_dfq_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
# classified on the opcode token alone (operands such as R2.64 must not count as a width)
MNEMONICS = [("LDG.256", r"^LDG\..*\b256\b"), ("LDG.128", r"^LDG\..*\b128\b"), ("LDG.64", r"^LDG\..*\b64\b"),
             ("LDG.32", r"^LDG\b"), ("STG.256", r"^STG\..*\b256\b"), ("STG.128", r"^STG\..*\b128\b"),
             ("STG.64", r"^STG\..*\b64\b"), ("STG.32", r"^STG\b"), ("LDS", r"^LDS"), ("STS", r"^STS"),
             ("UBLKCP (TMA bulk)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("RED/ATOM", r"^(REDG|RED|ATOMG|ATOM|ATOMS)\b"),
             ("BAR", r"^BAR\b"), ("MUFU", r"^MUFU")]


def demangle(names):
    out = subprocess.run(["/usr/local/cuda/bin/cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    # drop the trailing parameter list, keep the template arguments
    return [re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", n).replace("oodfq::", "").replace("void ", "") for n in out]


def main():
    print("# static report: ptxas -v and SASS instruction mix per kernel (sm_100a), " + " ".join(FLAGS))
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        with tempfile.TemporaryDirectory() as tmp:
            obj = os.path.join(tmp, "a.o")
            p = subprocess.run([NVCC] + FLAGS + ["-Xptxas", "-v", "-c", "-o", obj, src], capture_output=True, text=True)
            if p.returncode != 0:
                sys.exit(p.stderr)
            info, cur = {}, None
            for ln in p.stderr.splitlines():
                m = re.search(r"Compiling entry function '(\S+)'", ln)
                if m:
                    cur = m.group(1)
                    info[cur] = {"regs": "?", "smem": "0", "spill": "0"}
                elif cur and "Used" in ln:
                    info[cur]["regs"] = re.search(r"Used (\d+) registers", ln).group(1)
                    sm = re.search(r"(\d+) bytes smem", ln)
                    info[cur]["smem"] = sm.group(1) if sm else "0"
                elif cur and "spill stores" in ln:
                    info[cur]["spill"] = re.search(r"(\d+) bytes spill stores", ln).group(1)
            sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
            mix, fn = defaultdict(Counter), None
            for ln in sass.splitlines():
                m = re.search(r"Function : (\S+)", ln)
                if m:
                    fn = m.group(1)
                    continue
                m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if fn and m:
                    for label, pat in MNEMONICS:
                        if re.search(pat, m.group(1)):
                            mix[fn][label] += 1
                            break
        if not info:
            continue
        print(f"\n## {os.path.basename(src)}")
        names = list(info)
        for mangled, nice in zip(names, demangle(names)):
            i = info[mangled]
            ops = ", ".join(f"{k} x{v}" for k, v in mix.get(mangled, {}).items())
            print(f"{nice[:96]:96s} regs {i['regs']:>3s}  smem {i['smem']:>6s} B  spill {i['spill']} B  | {ops}")


if __name__ == "__main__":
    main()
