#!/usr/bin/env python
"""Check that every kernel of a validated build is instruction-identical in the current library.

Used when code is added after the GPU budget of a round is spent: experimental variants are added as new
template instantiations / new kernels, and this script proves that the kernels that passed the GPU parity
suite were not perturbed.

  python scripts/check_sass_identity.py <validated commit> [--rename OLD=NEW ...]
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]


def sass_by_func(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is not None:
            t = re.sub(r"^/\*[0-9a-f]{4}\*/ ", "", re.sub(r"\s+", " ", line).strip())
            # keep instruction lines (".. ; /* 0x<encoding> */") and their second encoding word only:
            # fatbin section headers that follow the last function of a file are not code
            if re.search(r"/\* 0x[0-9a-f]{16} \*/$", t):
                funcs[cur].append(t)
    return funcs


def main():
    commit = sys.argv[1]
    renames = dict(a.split("=", 1) for a in sys.argv[3:]) if len(sys.argv) > 2 and sys.argv[2] == "--rename" else {}
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "uni3detr_b200/csrc", "include"],
                             capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        srcs = sorted(os.path.join(tmp, "uni3detr_b200/csrc", f) for f in os.listdir(os.path.join(tmp, "uni3detr_b200/csrc"))
                      if f.endswith(".cu"))
        so = os.path.join(tmp, "validated.so")
        subprocess.run(["nvcc"] + FLAGS + ["-o", so] + srcs, check=True, cwd=tmp)
        old = sass_by_func(so)
    new = sass_by_func(os.path.join(ROOT, "uni3detr_b200", "libu3d_b200.so"))
    bad = []
    for name, code in old.items():
        n = name
        for a, b in renames.items():
            n = n.replace(a, b)
        if n not in new:
            bad.append((name, "missing"))
        elif new[n] != code:
            bad.append((name, "changed"))
    print(f"{len(old)} validated kernels, {len(new)} in the current library, {len(new) - len(old)} new, "
          f"{len(bad)} missing/changed")
    for b in bad:
        print("  ", b)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
