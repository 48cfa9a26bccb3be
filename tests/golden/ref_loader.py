"""Import the unmodified reference from ``/root/reference/src`` (this container only).

The reference loads ``core/data/SimpleOpt1-5/ALL.bbdep.rotamers.lib`` at import
(``mcsce/core/rotamer_library.py:16`` via ``side_chain_builder.py:27``); that
archive is missing from the checkout and ``/root/reference`` is read-only, so the
module-level path is pointed at the seeded synthetic library before
``side_chain_builder`` is imported.  Nothing else in the reference is touched.

Used only by ``make_golden.py`` and by tests that are skipped when the reference
is absent (it does not exist on the GPU box).
"""
import os
import sys
import tempfile
from pathlib import Path

REF_SRC = Path(os.environ.get("MCSCE_REFERENCE_SRC", "/root/reference/src"))


def reference_available():
    return (REF_SRC / "mcsce" / "core" / "side_chain_builder.py").exists()


def load_reference(lib_seed=0):
    """Return the reference's ``mcsce`` package with the synthetic library wired in."""
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    repo = Path(__file__).resolve().parents[2]
    if str(repo) not in sys.path:
        sys.path.insert(0, str(repo))
    from mcsce_b200 import synth

    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_ref"))
    sys.dont_write_bytecode = True
    if str(REF_SRC) not in sys.path:
        sys.path.insert(0, str(REF_SRC))
    import mcsce.core.rotamer_library as rl

    rl._library_path = synth.synthetic_library_path(lib_seed)
    import mcsce
    import mcsce.core.side_chain_builder  # noqa: F401  (instantiates the library)
    return mcsce
