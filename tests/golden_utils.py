"""Helpers shared by the parity tests: load a golden fixture and replay it."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


class GoldenCase:
    def __init__(self, name):
        self.name = name
        self.z = np.load(GOLDEN / (name + ".npz"), allow_pickle=False)
        self.pdb = str(self.z["pdb"])
        self.terms = tuple(str(t) for t in self.z["terms"])
        self.beta = float(self.z["beta"])
        self.n_conf = int(self.z["n_conf"])
        self.full = bool(self.z["full"])
        # residue numbers that keep the side chains of the input structure (--fix mode, cli.py:118-147)
        self.retain = tuple(int(r) for r in self.z["retain"]) if "retain" in self.z.files else ()

    def structure(self):
        from mcsce_b200.structure import Structure

        s = Structure(self.pdb).build()
        return s.remove_side_chains(list(self.retain))

    def n_steps(self, c):
        return int(self.z["c%d_nsteps" % c])

    def step(self, c, k):
        pre = "c%d_s%d_" % (c, k)
        d = {key[len(pre):]: self.z[key] for key in self.z.files if key.startswith(pre)}
        return d

    def conf(self, c):
        pre = "c%d_" % c
        return {key[len(pre):]: self.z[key] for key in self.z.files
                if key.startswith(pre) and not key[len(pre):].startswith("s")}

    def steps_by_level(self, c):
        """{residue ordinal: step record}; the reference's energy-closure level is ordinal+1."""
        return {int(self.step(c, k)["level"]) - 1: self.step(c, k) for k in range(self.n_steps(c))}


def library():
    from mcsce_b200.core.rotamer_library import DunbrakRotamerLibrary, ptmRotamerLib

    global _LIB
    try:
        return _LIB
    except NameError:
        from mcsce_b200 import synth

        # the fixtures were generated on the seeded synthetic library (the Dunbrack archive is not in the reference checkout)
        _LIB = (DunbrakRotamerLibrary(path=synth.synthetic_library_path(0)), ptmRotamerLib())
        return _LIB
