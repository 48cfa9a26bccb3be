"""Randomised bundles (standard + PTM residues, 1-3 helices, 1-64 conformers, forced splits): the growth must not depend
on which of its shortcuts are on, and the C oracle replaying the GPU's chi rows and uniforms must agree with it."""
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


def test_random_systems_shortcuts_and_oracle_replay():
    import fuzz
    from oracle.growth_c import COracle

    trials = flips = 0
    for seed in range(500, 516):
        k = fuzz.make_case(seed)
        s, sysm, eng, C, split, u, sd = (k[x] for x in ("s", "sysm", "eng", "C", "split", "u", "sd"))
        a = eng.grow(C, fuzz.BETA, seed=sd, uniforms=u, dump=True, no_graph=True, env_split=split)
        a = {x: a[x].cpu().numpy().copy() for x in ("energy", "ok", "selected", "trial_energy", "chis")}
        # neighbour screen off / chi-independent atoms in every trial / CUDA-graph replay: same growth
        for kw in (dict(no_graph=True, no_screen=True), dict(no_graph=True, no_screen=True, no_dedup=True), dict()):
            b = eng.grow(C, fuzz.BETA, seed=sd, uniforms=u, dump=True, env_split=split, **kw)
            assert np.array_equal(a["selected"], b["selected"].cpu().numpy()), (seed, kw)
            assert np.array_equal(a["ok"], b["ok"].cpu().numpy()), (seed, kw)
            if "no_dedup" not in kw:
                assert np.array_equal(a["energy"], b["energy"].cpu().numpy(), equal_nan=True), (seed, kw)
        ref = COracle(sysm, s.coords).grow(C, fuzz.BETA, chis=a["chis"], uniforms=u, dump=True)
        te, tr = a["trial_energy"], ref["trial_energy"]
        both = ~np.isnan(te) & ~np.isnan(tr)
        trials += int(both.sum())
        flips += int((np.isinf(te[both]) != np.isinf(tr[both])).sum())
        fin = both & np.isfinite(te) & np.isfinite(tr)
        bad = np.abs(te[fin] - tr[fin]) > np.maximum(1e-3, 1e-5 * np.abs(tr[fin]))
        assert bad.mean() < 0.03, seed
    # K1's float32 differences (~1e-6 A) flip a clash decision only for a pair within ~2e-7 of its clash distance:
    # measured 9 flips in 2e7 trials (tools/fuzz.py, tools/fuzz_diag.py)
    assert trials > 100000 and flips <= 2, (trials, flips)
