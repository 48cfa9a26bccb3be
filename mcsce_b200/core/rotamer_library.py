"""Backbone-dependent rotamer library: the trial-set stage of the growth path.

Host-side Python, as the north-star keeps it.  Same entry points and results as
``mcsce/core/rotamer_library.py``: ``DunbrakRotamerLibrary`` (bbdep2010 text ->
``{(res, phi, psi): [chis (T, nchi) deg, probs (T,)]}`` with the 0.001
probability threshold and the +-sigma x3 / x9 augmentation, ``:52-129``),
``retrieve_torsion_and_prob`` with the PTM expansion (``:131-177``) and
``ptmRotamerLib`` (``:180-241``).  The loader is vectorised (pandas) instead of
line-by-line; the row order of every bin - which is the trial index the growth
loop selects by - is the reference's.

Added for the GPU path: :meth:`DunbrakRotamerLibrary.trial_distribution` returns
the per-trial chi means / sigmas / prior weights without drawing, so the device
can do the PTM ``N(val, sig)`` draw and the 0.5 deg perturbation itself.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from mcsce_b200.chem.tables import PTM_PARENT, PTM_ROTAMER_ALIAS, load_tables

_ONE_CHI = ("CYS", "SER", "THR", "VAL")
_ONE_CHI_AUGMENT = ("CYS", "SER", "VAL")      # THR is augmented 9-fold like the rest (rotamer_library.py:93)
_TWO_CHI = ("ASN", "ASP", "HIS", "ILE", "LEU", "PHE", "PRO", "TRP", "TYR")
_THREE_CHI = ("GLN", "GLU", "MET")


ALLOW_SYNTHETIC_ENV = "MCSCE_B200_ALLOW_SYNTHETIC"
_REAL_LIBRARY_NAMES = ("ALL.bbdep.rotamers.lib", "ALL.bbdep.rotamers.lib.gz")


def default_library_path():
    """bbdep library location, in this order: ``$MCSCE_ROTAMER_LIB``; the Dunbrack archive beside the package data
    (``mcsce_b200/data/SimpleOpt1-5/ALL.bbdep.rotamers.lib``, where the reference keeps it: ``core/rotamer_library.py:16``).
    The archive is absent from the reference checkout, so tests and the benchmark run on a *seeded synthetic* stand-in
    (random priors: scientifically meaningless ensembles) - that one is only served when the caller opts in with
    ``MCSCE_B200_ALLOW_SYNTHETIC=1`` (or passes ``synth.synthetic_library_path()`` explicitly); otherwise this raises."""
    env = os.environ.get("MCSCE_ROTAMER_LIB")
    if env:
        if not Path(env).exists():
            raise FileNotFoundError("MCSCE_ROTAMER_LIB=%s does not exist" % env)
        return Path(env)
    data = Path(__file__).resolve().parents[1] / "data"
    for sub in ("SimpleOpt1-5", ""):
        for name in _REAL_LIBRARY_NAMES:
            if (data / sub / name).exists():
                return data / sub / name
    if os.environ.get(ALLOW_SYNTHETIC_ENV, "") not in ("", "0"):
        import warnings

        from mcsce_b200 import synth

        warnings.warn("mcsce_b200: using the SYNTHETIC backbone-dependent rotamer library (seeded random priors); "
                      "ensembles grown from it are for testing/benchmarking only", stacklevel=2)
        return synth.synthetic_library_path(0)
    raise FileNotFoundError(
        "no backbone-dependent rotamer library: set MCSCE_ROTAMER_LIB to the Dunbrack bbdep file (the reference's "
        "data/SimpleOpt1-5/ALL.bbdep.rotamers.lib) or place it under %s; set %s=1 to run on the synthetic stand-in "
        "(tests and benchmarks only)" % (data / "SimpleOpt1-5", ALLOW_SYNTHETIC_ENV))


def wrap_angles(a):
    a = np.asarray(a, dtype=np.float64)
    a = np.where(a > 180.0, a - 360.0, a)
    return np.where(a < -180.0, a + 360.0, a)


def get_closest_angle(angle, degsep=10):
    """Nearest grid angle in [-180, 170]; Python ``round`` (banker's) as in the reference (:27-34)."""
    if angle > 180 - degsep / 2.0:
        return -180
    return round(angle / degsep) * degsep


def get_floor_angle(angle, degsep=120):
    if angle == 360.0:
        return 0
    return np.floor(angle / degsep) * degsep


class ptmRotamerLib:
    """SIDEpro-format PTM rotamers; served from the packed chemistry tables."""

    def __init__(self, probability_threshold=0.001):
        t = load_tables()
        self._info = t.ptm_info
        self._data = {k: v.tolist() for k, v in t.ptm_data.items()}

    def get_dependence(self, residue_type):
        return self._info[residue_type]

    def retrieve_torsion_and_prob(self, residue_type, dchi=-360.0):
        if self._info[residue_type][-1] == 0:
            return self._data[(residue_type, -360.0)]
        if dchi < 0:
            dchi += 360.0
        return self._data[(residue_type, float(get_floor_angle(dchi, 120.0)))]


class DunbrakRotamerLibrary:
    def __init__(self, probability_threshold=0.001, augment_with_std=True, path=None):
        import pandas as pd

        self.prob_threshold = probability_threshold
        self.path = Path(path) if path is not None else default_library_path()
        df = pd.read_csv(self.path, sep=r"\s+", comment="#", header=None,
                         names=["res", "phi", "psi", "cnt", "r1", "r2", "r3", "r4", "prob",
                                "c1", "c2", "c3", "c4", "s1", "s2", "s3", "s4"])
        df = df[(df.prob > probability_threshold)]
        have_m180 = set(df.res[(df.phi == -180) & (df.psi == -180)].unique())
        edge = df[(df.phi == 180) | (df.psi == 180)]
        missing = set(edge.res.unique()) - have_m180
        assert not missing, "library lists phi/psi=180 rows before the -180 rows for %s" % sorted(missing)
        df = df[(df.phi != 180) & (df.psi != 180)]
        res = df.res.to_numpy()
        phi = df.phi.to_numpy(); psi = df.psi.to_numpy()
        prob = df.prob.to_numpy(dtype=np.float64)
        chi = df[["c1", "c2", "c3", "c4"]].to_numpy(dtype=np.float64)
        sig = df[["s1", "s2"]].to_numpy(dtype=np.float64)
        if augment_with_std:
            one = np.isin(res, _ONE_CHI_AUGMENT)
            mult = np.where(one, 3, 9)
            newp = prob / mult
            keep = newp > probability_threshold
            res, phi, psi, chi, sig, mult, newp = (a[keep] for a in (res, phi, psi, chi, sig, mult, newp))
            src = np.repeat(np.arange(len(res)), mult)
            first = np.concatenate([[0], np.cumsum(mult)[:-1]])
            k = np.arange(len(src)) - np.repeat(first, mult)        # 0..mult-1 within each source row
            one_r = (mult == 3)[src]
            d1 = np.where(one_r, k, k // 3) - 1                     # -1, 0, +1 on chi1 (outer loop)
            d2 = np.where(one_r, 0, k % 3 - 1)                      # -1, 0, +1 on chi2 (inner loop)
            chis = chi[src].copy()
            # same arithmetic as the reference: chi - sigma, chi, chi + sigma
            chis[:, 0] = np.where(d1 < 0, chi[src, 0] - sig[src, 0], np.where(d1 > 0, chi[src, 0] + sig[src, 0], chi[src, 0]))
            chis[:, 1] = np.where(d2 < 0, chi[src, 1] - sig[src, 1], np.where(d2 > 0, chi[src, 1] + sig[src, 1], chi[src, 1]))
            probs = newp[src]
            res, phi, psi = res[src], phi[src], psi[src]
        else:
            chis, probs = chi, prob
        self._data = {}
        if len(res):
            key_change = np.concatenate([[True], (res[1:] != res[:-1]) | (phi[1:] != phi[:-1]) | (psi[1:] != psi[:-1])])
            starts = np.where(key_change)[0]
            ends = np.concatenate([starts[1:], [len(res)]])
            for s, e in zip(starts, ends):
                label = (str(res[s]), int(phi[s]), int(psi[s]))
                r = label[0]
                n = 1 if r in _ONE_CHI else 2 if r in _TWO_CHI else 3 if r in _THREE_CHI else 4
                block = [wrap_angles(chis[s:e, :n]), probs[s:e].copy()]
                if label in self._data:      # non-contiguous repeat of a bin: append in file order
                    block = [np.concatenate([self._data[label][0], block[0]]),
                             np.concatenate([self._data[label][1], block[1]])]
                self._data[label] = block

    # -- lookups --------------------------------------------------------------
    @staticmethod
    def _bin(phi, psi):
        if np.isnan(phi):
            phi = -180     # first residue has no phi (:132-134)
        if np.isnan(psi):
            psi = -180     # last residue has no psi (:135-137)
        return get_closest_angle(phi), get_closest_angle(psi)

    def retrieve_torsion_and_prob(self, residue_type, phi, psi, ptmlib):
        """[chis (T, nchi) degrees, probs (T,)]; PTM extra chis are drawn with the global
        NumPy RNG exactly where the reference draws them (:139-172)."""
        bphi, bpsi = self._bin(phi, psi)
        if residue_type in PTM_PARENT:
            chis, probs = self._data[(PTM_PARENT[residue_type], bphi, bpsi)]
            residue_type = PTM_ROTAMER_ALIAS.get(residue_type, residue_type)
            if residue_type not in ptmlib._info:
                print("ptm residue rotamers not provided, assumes rotamers of unmodified residue")
                return [chis, probs]
            n_total, n_extra, dep = ptmlib.get_dependence(residue_type)
            if dep == 0:
                rows = np.array(ptmlib.retrieve_torsion_and_prob(residue_type, -360.0))
                tors = wrap_angles(np.random.normal(rows[:, 1:1 + n_extra], rows[:, 1 + n_extra:]))
                assert tors.shape[1] == n_total
                return [tors, rows[:, 0]]
            new_chis, new_probs = [], []
            for i in range(chis.shape[0]):
                rows = np.array(ptmlib.retrieve_torsion_and_prob(residue_type, chis[i, dep - 1]))
                tors = wrap_angles(np.random.normal(rows[:, 1:1 + n_extra], rows[:, 1 + n_extra:]))
                for j in range(len(rows)):
                    p = probs[i] * rows[j, 0]
                    if p > self.prob_threshold:
                        new_chis.append(np.concatenate((chis[i], tors[j])))
                        new_probs.append(p)
            assert len(new_chis[0]) == n_total
            return [np.array(new_chis), np.array(new_probs)]
        if residue_type in ("HID", "HIE", "HIP"):
            residue_type = "HIS"
        return self._data[(residue_type, bphi, bpsi)]

    def trial_distribution(self, residue_type, phi, psi, ptmlib):
        """(mean (T, nchi), sigma (T, nchi), prob (T,)) of the trial set without drawing.

        sigma is zero for library chis and the PTM table's sigma for PTM extra chis; a
        draw of the reference's ``retrieve_torsion_and_prob`` is
        ``wrap(mean + sigma * z)`` with independent standard normals z per trial and chi
        (each element of the reference's per-parent draw is used by exactly one trial).
        """
        bphi, bpsi = self._bin(phi, psi)
        if residue_type in PTM_PARENT:
            chis, probs = self._data[(PTM_PARENT[residue_type], bphi, bpsi)]
            alias = PTM_ROTAMER_ALIAS.get(residue_type, residue_type)
            if alias not in ptmlib._info:
                return chis.copy(), np.zeros_like(chis), probs.copy()
            n_total, n_extra, dep = ptmlib.get_dependence(alias)
            if dep == 0:
                rows = np.array(ptmlib.retrieve_torsion_and_prob(alias, -360.0))
                return rows[:, 1:1 + n_extra].copy(), rows[:, 1 + n_extra:].copy(), rows[:, 0].copy()
            mean, sigma, out_p = [], [], []
            for i in range(chis.shape[0]):
                rows = np.array(ptmlib.retrieve_torsion_and_prob(alias, chis[i, dep - 1]))
                for j in range(len(rows)):
                    p = probs[i] * rows[j, 0]
                    if p > self.prob_threshold:
                        mean.append(np.concatenate((chis[i], rows[j, 1:1 + n_extra])))
                        sigma.append(np.concatenate((np.zeros(chis.shape[1]), rows[j, 1 + n_extra:])))
                        out_p.append(p)
            return np.array(mean), np.array(sigma), np.array(out_p)
        if residue_type in ("HID", "HIE", "HIP"):
            residue_type = "HIS"
        chis, probs = self._data[(residue_type, bphi, bpsi)]
        return chis.copy(), np.zeros_like(chis), probs.copy()
