"""Recipe: stage the reference's own Python modules of the hot path under oracle/_ref/ (git-ignored, travels to the GPU box).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python, so "building" it = copying the two packages
its forward and decode chain live in - ``model/`` (MESM, transformer, attention, position encoding, criterion) and
``utils/`` (span_utils, temporal_nms, post_processing, data_utils) - unmodified, from where they lie under
/root/reference.  Nothing is copied into the repository history: oracle/_ref/ is listed in .gitignore (not in
.gpurunignore).  bench.py's reference legs import these modules (``cpu_baseline.kind = "reference"``, and the same
modules in PyTorch eager mode on the B200 as the GPU bar); without them they fall back to the oracle port (``"port"``).

    python -m oracle.make_ref            # run by __graft_entry__.build() whenever /root/reference is present
"""
import os
import shutil
import sys

REF = os.environ.get("MESM_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGES = ("model", "utils")


def make_ref(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/make_ref: {REF} not present; keeping whatever oracle/_ref already holds")
        return os.path.isdir(os.path.join(DST, "model"))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for pkg in PACKAGES:
        src, dst = os.path.join(REF, pkg), os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(dst)
        for f in sorted(os.listdir(src)):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
                n += 1
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copies of lntzm/MESM model/*.py and utils/*.py staged by oracle/make_ref.py (git-ignored; "
                "test infrastructure for the reference baselines of bench.py).\n")
    if verbose:
        print(f"oracle/make_ref: staged {n} files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
