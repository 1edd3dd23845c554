"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the pure-Python reference modules of the hot path.
TEST INFRASTRUCTURE ONLY.

/root/reference exists in the authoring container but not on the GPU box.  The reference is pure Python
(no build step), so "building" it means copying the module files that `cap4d.mmdm` needs at import time -
nothing is edited - into oracle/_ref/ (listed in .gitignore, NOT in .gpurunignore: it travels to the GPU
box like the built .so files and never enters the history).  With it the unmodified `MMLDM`,
`MMDMUnetModel` and `StochasticIOSampler` run on the box:

  * tests/test_gpu_dropin.py drives the reference's own sampler and `apply_model` over `install(mmldm)`;
  * `bench.py --impl reference` times the reference U-Net itself (cpu_baseline.kind = "reference").

    python oracle/build_ref.py            (also run by __graft_entry__.build())
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("CAP4D_REFERENCE_ROOT", "/root/reference")

# package directories whose *.py files are copied (everything `import cap4d.mmdm.mmdm` / `.sampler` reaches,
# plus the conditioning module the conditioning oracle is pinned against)
PACKAGES = ["cap4d/mmdm", "controlnet/ldm", "controlnet/cldm"]
TOP_LEVEL = ["cap4d/__init__.py", "controlnet/__init__.py"]


def build(verbose: bool = True) -> str:
    """Copy the reference modules; returns DEST, or '' when the reference is not present (GPU box: the
    prebuilt copy is used as it is)."""
    if not os.path.isdir(os.path.join(SRC, "cap4d", "mmdm")):
        if verbose:
            print(f"oracle/_ref: {SRC} not present, keeping the existing copy" if os.path.isdir(DEST) else
                  f"oracle/_ref: {SRC} not present and no copy exists")
        return DEST if os.path.isdir(DEST) else ""
    n = 0
    for pkg in PACKAGES:
        root = os.path.join(SRC, pkg)
        if not os.path.isdir(root):
            continue
        for d, _, files in os.walk(root):
            rel = os.path.relpath(d, SRC)
            for f in files:
                if not f.endswith(".py"):
                    continue
                os.makedirs(os.path.join(DEST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), os.path.join(DEST, rel, f))
                n += 1
    for f in TOP_LEVEL:
        src = os.path.join(SRC, f)
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif not os.path.exists(dst):
            open(dst, "w").close()  # namespace marker only
    with open(os.path.join(DEST, "README"), "w") as fh:
        fh.write("Verbatim copy of reference modules made by oracle/build_ref.py; git-ignored; do not edit.\n")
    if verbose:
        print(f"oracle/_ref: copied {n} reference modules from {SRC}")
    return DEST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
