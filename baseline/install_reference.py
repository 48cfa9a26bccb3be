"""Install the UNMODIFIED reference into ``baseline/_ref`` (git-ignored, travels to the GPU box).

    python baseline/install_reference.py

The reference's ``setup.py`` unzips ``src/mcsce/core/data/SimpleOpt1-5.zip`` (the Dunbrack library), a large blob that is
missing from the checkout (``/root/reference/.MISSING_LARGE_BLOBS``).  ``/root/reference`` is read-only, so the tree is
copied to /tmp, the missing archive is *supplied* there - a zip holding the seeded synthetic library of
``mcsce_b200.synth`` under the file name the reference loads (``core/rotamer_library.py:16``) - and the copy is
installed with pip (offline).  No reference source file is edited.
"""
import shutil
import subprocess
import sys
import tempfile
import zipfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
TARGET = ROOT / "baseline" / "_ref"


def main():
    from mcsce_b200 import synth

    if not (REF / "setup.py").exists():
        raise SystemExit("reference checkout not present at %s" % REF)
    tmp = Path(tempfile.mkdtemp(prefix="mcsce_ref_"))
    src = tmp / "reference"
    shutil.copytree(REF, src)
    lib = Path(synth.synthetic_library_path(0))
    with zipfile.ZipFile(src / "src" / "mcsce" / "core" / "data" / "SimpleOpt1-5.zip", "w", zipfile.ZIP_DEFLATED) as z:
        z.write(lib, "ALL.bbdep.rotamers.lib")
    if TARGET.exists():
        shutil.rmtree(TARGET)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", str(TARGET), str(src)]
    print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    data = TARGET / "mcsce" / "core" / "data" / "SimpleOpt1-5"
    if not (data / "ALL.bbdep.rotamers.lib").exists():          # package_data may not carry the unzipped file: place it
        data.mkdir(parents=True, exist_ok=True)
        shutil.copy(lib, data / "ALL.bbdep.rotamers.lib")
    print("installed:", TARGET)


if __name__ == "__main__":
    main()
