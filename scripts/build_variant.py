#!/usr/bin/env python
"""Build a variant of libpadt_b200.so for same-box A/B measurements (boxes differ by ~10 %, so only A/B inside ONE gpurun call
is meaningful).   python scripts/build_variant.py <name> [--rev GITREV] [-DMACRO=V ...]   ->  padt_b200/variants/lib_<name>.so"""
import glob, os, subprocess, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
rev = None
defs = []
args = sys.argv[2:]
while args:
    a = args.pop(0)
    if a == "--rev":
        rev = args.pop(0)
    else:
        defs.append(a)
src_root = ROOT
tmp = None
if rev:
    tmp = tempfile.mkdtemp()
    subprocess.check_call(f"git -C {ROOT} archive {rev} padt_b200/csrc include | tar -x -C {tmp}", shell=True)
    src_root = tmp
csrc = os.path.join(src_root, "padt_b200", "csrc")
out_dir = os.path.join(ROOT, "padt_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
bdir = tempfile.mkdtemp()
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "177,550"] + defs
procs, objs = [], []
for src in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
    obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen(["nvcc"] + flags + ["-c", src, "-o", obj]))
for p in procs:
    if p.wait() != 0:
        sys.exit("nvcc failed")
out = os.path.join(out_dir, f"lib_{name}.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", out] + objs)
shutil.rmtree(bdir)
if tmp:
    shutil.rmtree(tmp)
print("built", out)
