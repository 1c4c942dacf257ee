"""Places the UNMODIFIED reference package under baseline/_ref (git-ignored, travels to the GPU box).

`pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` fails in this image: the reference's build backend (poetry-core, pyproject.toml) is
not in the wheelhouse ("ModuleNotFoundError: No module named 'poetry'").  What that command would have
produced for this pure-Python package is the package directory plus a dist-info directory; this script
writes exactly those two: a byte-for-byte copy of /root/reference/turtlemd and a minimal
turtlemd-<version>.dist-info/METADATA (the reference asks importlib.metadata for its own version at
import time, turtlemd/version.py:4).  No reference source enters the git history.

    python baseline/install_ref.py [/root/reference]
"""
import pathlib
import re
import shutil
import sys

HERE = pathlib.Path(__file__).resolve().parent


def install(reference="/root/reference"):
    src = pathlib.Path(reference)
    if not (src / "turtlemd").is_dir():
        return None
    dst = HERE / "_ref"
    if dst.exists():
        shutil.rmtree(dst)
    dst.mkdir(parents=True)
    shutil.copytree(src / "turtlemd", dst / "turtlemd", ignore=shutil.ignore_patterns("__pycache__"))
    version = re.search(r'^version\s*=\s*"([^"]+)"', (src / "pyproject.toml").read_text(), re.M).group(1)
    info = dst / f"turtlemd-{version}.dist-info"
    info.mkdir()
    (info / "METADATA").write_text(f"Metadata-Version: 2.1\nName: turtlemd\nVersion: {version}\n")
    (info / "INSTALLER").write_text("baseline/install_ref.py\n")
    (info / "RECORD").write_text("")
    return dst


if __name__ == "__main__":
    out = install(*sys.argv[1:2])
    print(f"reference installed under {out}" if out else "no reference tree found: nothing installed")
