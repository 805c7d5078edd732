"""Recipe for oracle/_ref/: the UNMODIFIED reference package files the hot path lives in, so that the GPU box
(which has no /root/reference) can time the reference's own Numba kernels as the CPU baseline.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python (no build step): "building" it is copying
rag_system/core/*.py from where they lie under /root/reference into the git-ignored oracle/_ref/ tree
(never into the repository's history).  __graft_entry__.build() runs this when /root/reference is present;
oracle/ref_runner.py imports the result when it exists, and bench.py then reports cpu_baseline.kind = "reference"
instead of "port"."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2R_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make(verbose: bool = True) -> bool:
    src = os.path.join(REF, "rag_system", "core")
    if not os.path.isdir(src):
        if verbose:
            print(f"[make_ref] {src} not present: oracle/_ref left as it is")
        return False
    dst = os.path.join(DST, "rag_system", "core")
    os.makedirs(dst, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
            n += 1
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write(f"verbatim copies of {src}/*.py ({n} files), made by oracle/make_ref.py; git-ignored\n")
    if verbose:
        print(f"[make_ref] {n} reference files -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
