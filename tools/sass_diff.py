#!/usr/bin/env python
"""Which kernels of the working tree differ, instruction for instruction, from the build that last ran on a B200?

    python tools/sass_diff.py [--base 16ddc86] [--out profiles/r02_sass_diff.txt]

Builds the objects of `--base` (default: the last commit whose library executed on hardware in round 1) from `git show` into a temporary
directory and the working tree's objects next to them (nvcc, sm_100a, the flags of __graft_entry__.build), dumps both with
`cuobjdump -sass`, strips instruction encodings, constant-bank offsets and register numbers (a new ChainArgs field moves every
c[0x0][..] offset; register allocation is not stable across unrelated edits) and compares kernel by kernel.  A kernel listed as
identical is the binary round 1 measured; everything else is what tools/preflight.py and the probes exist for."""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
SOURCES = ["api.cu", "softmax_net.cu"]


def build_tree(src_root: str, out_dir: str):
    os.makedirs(out_dir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(out_dir, s.replace(".cu", ".o"))
        procs.append((obj, subprocess.Popen(["nvcc", *FLAGS, "-c", "-o", obj, os.path.join(src_root, "sgmcmc_b200", "csrc", s)], cwd=src_root,
                                            stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
    for obj, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise SystemExit(f"nvcc failed for {obj}:\n{err[-2000:]}")
    return [obj for obj, _ in procs]


def kernels(objs):
    """demangled kernel name -> normalised SASS lines"""
    out = {}
    for obj in objs:
        text = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
        name, body = None, []
        for ln in text.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                if name:
                    out[name] = body
                name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
                name = re.sub(r"\(.*$", "", name)     # argument lists change with unrelated struct edits
                body = []
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
            if m and name:
                ins = m.group(1)
                ins = re.sub(r"c\[0x[0-9a-f]+\]\[0x[0-9a-f]+\]", "c[][]", ins)
                ins = re.sub(r"\b(U?R|U?P)\d+\b", r"\1", ins)
                ins = re.sub(r"0x[0-9a-f]+", "0x", ins)   # immediates that are addresses / offsets
                body.append(ins.strip())
        if name:
            out[name] = body
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--base", default="16ddc86")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_diff.txt"))
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        base_src = os.path.join(tmp, "base")
        os.makedirs(base_src)
        subprocess.check_call(f"git archive {args.base} sgmcmc_b200/csrc include | tar -x -C {base_src}", shell=True, cwd=ROOT)
        base = kernels(build_tree(base_src, os.path.join(tmp, "obj_base")))
        head = kernels(build_tree(ROOT, os.path.join(tmp, "obj_head")))
    same = [k for k in head if k in base and head[k] == base[k]]
    changed = [k for k in head if k in base and head[k] != base[k]]
    new = [k for k in head if k not in base]
    gone = [k for k in base if k not in head]
    rev = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True, cwd=ROOT).strip()
    lines = [f"Round 2 -- which kernels of the working tree (HEAD {rev}) differ from the build that ran on a B200 in round 1 (commit {args.base}, the last one",
             "whose library executed on hardware).  Method (tools/sass_diff.py): cuobjdump -sass of both builds' objects, per kernel, with instruction",
             "encodings, constant-bank offsets, immediates and register numbers stripped.  Numbers: SASS instructions HEAD / round 1.", "",
             f"identical {len(same)} changed {len(changed)} new {len(new)} gone {len(gone)}"]
    lines += [f"CHANGED {k} {len(head[k])} {len(base[k])}" for k in sorted(changed)]
    lines += [f"NEW {k} {len(head[k])}" for k in sorted(new)]
    lines += [f"GONE {k} {len(base[k])}" for k in sorted(gone)]
    lines += [f"IDENTICAL {k} {len(head[k])}" for k in sorted(same)]
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:6]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
