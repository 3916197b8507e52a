#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE.  Stages the reference's own PyTorch flow module (src/flows/*.py, pure torch / numpy /
scipy, ~400 lines) as oracle/_ref/flows/ so that `bench.py --impl reference` can time the UNMODIFIED reference
implementation of the path (NormalizingFlowModel.forward, src/flows/models.py:11-24) on the GPU box's host cores.

    python oracle/make_ref.py [/root/reference]

The reference is Python: "building" it is a verbatim copy of the package from where it lies under the reference
checkout.  oracle/_ref/ is git-ignored (no reference source enters the history) but travels to the GPU box with the
repo snapshot, like the built .so files.  /root/reference only exists in the build container: nothing reads it at run
time.  Run by __graft_entry__.build() when the reference checkout is present."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("__init__.py", "flows.py", "models.py", "prior_dist.py", "utils.py")


def main(ref_root="/root/reference"):
    src = os.path.join(ref_root, "src", "flows")
    if not os.path.isdir(src):
        print(f"make_ref: {src} not present, nothing staged (the bench then times the C port of oracle/)")
        return 1
    dst = os.path.join(HERE, "_ref", "flows")
    os.makedirs(dst, exist_ok=True)
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    with open(os.path.join(HERE, "_ref", "STAGED_FROM"), "w") as fh:
        fh.write(f"{src}\nfiles: {' '.join(FILES)}\n")
    print(f"make_ref: staged {len(FILES)} files into {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:2]))
