#!/usr/bin/env python3
"""Recipe for oracle/_ref: a VERBATIM copy of the reference's package directory (pure Python, eight files), made
where the checkout is mounted so that the unmodified reference can run on the GPU box too (tests/
test_oracle_vs_reference.py, `bench.py --impl reference`, `cpu_baseline.kind = "reference"`).

    python oracle/make_ref.py [/root/reference]          (or: make -C oracle)

oracle/_ref/ is build output: git-ignored (reference sources never enter the history), not gpurun-ignored.
Nothing is compiled -- the reference has no native code -- and nothing is edited: the copy is checked byte for byte.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def make_ref(ref_root="/root/reference", quiet=False):
    src = os.path.join(ref_root, "gonzag")
    if not os.path.isfile(os.path.join(src, "bilin_mapping.py")):
        if not quiet:
            print("make_ref: no reference checkout at %s -- keeping whatever oracle/_ref holds" % ref_root)
        return None
    dst = os.path.join(HERE, "_ref", "gonzag")
    os.makedirs(dst, exist_ok=True)
    names = sorted(n for n in os.listdir(src) if n.endswith(".py"))
    for n in names:
        shutil.copyfile(os.path.join(src, n), os.path.join(dst, n))
    bad = [n for n in names if not filecmp.cmp(os.path.join(src, n), os.path.join(dst, n), shallow=False)]
    if bad:
        raise RuntimeError("copy differs from the reference: %s" % bad)
    with open(os.path.join(HERE, "_ref", "README"), "w") as f:
        f.write("Verbatim copy of %s/*.py made by oracle/make_ref.py (build output, git-ignored).\n" % src)
    if not quiet:
        print("oracle/_ref/gonzag: %d files copied from %s" % (len(names), src))
    return dst


if __name__ == "__main__":
    make_ref(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
