"""Stages the reference's own hot-path Python files into ``oracle/_ref/`` so they can travel to the GPU box.

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``oracle/_ref/`` is git-ignored (reference
sources never enter this repository's history) but NOT gpurun-ignored, so the staged copy ships with the snapshot like the built
``.so`` files do.  Staged: ``src/PaDT/models/{padt.py, padt_decoder.py, padt_processor.py}`` -- the three files that hold the
generate/decode path -- byte for byte, plus a manifest with their SHA-256.

Used by: ``__graft_entry__.build()`` (stages when /root/reference is present), ``oracle/ref_shim.py`` (imports the staged copy when
/root/reference is absent), ``bench.py --impl reference`` and the ``hf_gpu_baseline`` leg (the UNMODIFIED reference on the box's CPU
cores / on the same B200), ``tests/test_reference_floor_gpu.py``.

    python oracle/stage_ref.py            # -> oracle/_ref/PaDT/models/*.py + oracle/_ref/MANIFEST.json
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
FILES = ("padt.py", "padt_decoder.py", "padt_processor.py")
DST = os.path.join(HERE, "_ref")


def staged_src():
    """Directory to put on sys.path so that ``PaDT.models.*`` resolves to the staged files, or None."""
    d = os.path.join(DST, "PaDT", "models")
    return DST if all(os.path.exists(os.path.join(d, f)) for f in FILES) else None


def stage(verbose=True):
    src_dir = os.path.join(REF_ROOT, "src", "PaDT", "models")
    if not os.path.isdir(src_dir):
        return staged_src()
    dst_dir = os.path.join(DST, "PaDT", "models")
    os.makedirs(dst_dir, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(dst_dir, f))
        with open(os.path.join(dst_dir, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    head = os.path.join(REF_ROOT, ".git", "HEAD")
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_dir, "sha256": manifest, "note": "byte-for-byte copies; git-ignored; see oracle/stage_ref.py"}, fh, indent=1)
    if verbose:
        print(f"[stage_ref] staged {len(FILES)} reference files into {dst_dir}")
    return DST


if __name__ == "__main__":
    stage()
