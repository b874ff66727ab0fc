"""Install the UNMODIFIED reference into the git-ignored baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py            # in the build container (needs /root/reference)

The reference (SkylerGao/MC_NeRF) is not a pip package (no setup.py / pyproject.toml), so "install" = a plain copy of
its Python sources and yaml config, byte for byte: main.py, config/, data/, model/, utils/.  Nothing is edited; the six
packages it imports that are absent from this image (lpips, apriltag, prettytable, matplotlib(.pyplot/.cm),
mpl_toolkits.mplot3d) are replaced at import time by the inert stand-ins of baseline/ref_loader.py.

baseline/_ref/ is listed in .gitignore (reference sources never enter this repository's history) but not in
.gpurunignore, so `bench.py --impl reference` and tests/test_main_integration_gpu.py find it on the GPU box.
__graft_entry__.build() runs this whenever /root/reference is present.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yaml", ".txt", ".md")
TOP = ("main.py", "config", "data", "model", "utils", "LICENSE", "README.md")


def install(src=SRC, dst=DST, verbose=True):
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not present: the reference can only be installed in the build container")
    n = 0
    for top in TOP:
        s = os.path.join(src, top)
        if os.path.isfile(s):
            os.makedirs(dst, exist_ok=True)
            shutil.copyfile(s, os.path.join(dst, top))
            n += 1
            continue
        for root, dirs, files in os.walk(s):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                rel = os.path.relpath(os.path.join(root, f), src)
                out = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), out)
                n += 1
    # verify byte identity of what the baseline arm executes
    for rel in ("main.py", "model/mc_nerf.py", "model/net_block.py", "model/net_utils.py", "model/loss.py"):
        assert filecmp.cmp(os.path.join(src, rel), os.path.join(dst, rel), shallow=False), rel
    if verbose:
        print(f"installed {n} reference files into {dst}")
    return dst


if __name__ == "__main__":
    install()
    sys.exit(0)
