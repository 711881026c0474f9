"""Recipe that stages the UNMODIFIED reference next to the repo for the CPU reference arm of bench.py.
TEST / MEASUREMENT INFRASTRUCTURE ONLY.

The reference is a plain Python project without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build; this script does what that install would do: it places the reference's own
packages (model/, data/, trainer/, utils/ - the files SURVEY.md §8a cites), byte for byte, under baseline/_ref/.
That directory is git-ignored (never committed, never redistributed) but NOT gpurun-ignored, so it travels with the
snapshot to the GPU box where /root/reference does not exist.  Nothing under the product package imports it;
bench.py --impl reference (and the cpu_baseline leg, which shells out to it) are its only users.

    python -m oracle.install_ref        # no-op when /root/reference is absent
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("model", "data", "trainer", "utils")


def install(force=False):
    """Returns the install directory, or None when the reference checkout is not available here."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(os.path.join(DST, "model")) else None
    os.makedirs(DST, exist_ok=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(SRC, pkg), os.path.join(DST, pkg)
        if not os.path.isdir(src):
            continue
        if os.path.isdir(dst):
            if not force:
                continue
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return DST


def available():
    return os.path.isfile(os.path.join(DST, "model", "conformer.py")) and os.path.isfile(os.path.join(DST, "trainer", "trainer.py"))


if __name__ == "__main__":
    print(install(force=True))
