"""Deterministic synthetic inputs for the side-chain growth path.

Nothing ships with the reference for this path: the Dunbrack library archive is
missing from the checkout (``/root/reference/.MISSING_LARGE_BLOBS``) and there
are no sample PDBs.  This module generates (a) a backbone-dependent rotamer
library in the 17-column bbdep2010 text format that
``mcsce/core/rotamer_library.py:65-129`` parses and (b) backbone-only PDB text
with the atom set ``Structure.check_backbone_atom_completeness`` expects
(``mcsce/libs/libstructure.py:453-481``).  Both are seeded, so the golden
fixtures, the oracle, the GPU path and the benchmark all see the same inputs.
"""
from __future__ import annotations

import hashlib
import os
import tempfile
from pathlib import Path

import numpy as np

# ----------------------------------------------------------------------------
# synthetic bbdep rotamer library
# ----------------------------------------------------------------------------

#: residues present in a bbdep2010 library and the number of chi columns that
#: carry information for each (rotamer_library.py:118-127 trims to these).
LIB_RESIDUES = {
    "ARG": 4, "ASN": 2, "ASP": 2, "CYS": 1, "GLN": 3, "GLU": 3, "HIS": 2,
    "ILE": 2, "LEU": 2, "LYS": 4, "MET": 3, "PHE": 2, "PRO": 2, "SER": 1,
    "THR": 1, "TRP": 2, "TYR": 2, "VAL": 1,
}
_WELLS = np.array([-65.0, 180.0, 65.0])
_GRID = np.arange(-180, 181, 10)


#: prior over the three wells (-65, 180, +65): chi1 avoids g+ (it points back over the residue's own
#: backbone, as the real bbdep library knows for helical phi/psi), later chis favour trans
_CHI1_PRIOR = np.array([0.55, 0.37, 0.08])
_CHIN_PRIOR = np.array([0.28, 0.44, 0.28])


def write_synthetic_bbdep_library(path, seed=0, concentration=0.3, well_prior=True):
    """Write a seeded bbdep-format library to ``path``; returns the path.

    Rotamers are the 3**nchi combinations of three wells per chi; per
    (residue, phi, psi) bin the probabilities are Dirichlet draws (printed with
    6 decimals) whose concentration follows a per-well prior (``well_prior``:
    chi1 rarely g+, later chis mostly trans; total concentration
    ``concentration`` * 3**nchi), well centres are jittered by N(0, 4 deg) and
    sigmas are U(6, 14) deg.  The phi=180 / psi=180 rows repeat the -180
    rows and come after them, as the reference's loader asserts
    (rotamer_library.py:77-80).
    """
    rng = np.random.default_rng(seed)
    path = Path(path)
    chunks = [
        "# synthetic backbone-dependent rotamer library (mcsce_b200.synth, seed=%d)\n" % seed,
        "# T  Phi Psi Count r1 r2 r3 r4 Probabil chi1Val chi2Val chi3Val chi4Val chi1Sig chi2Sig chi3Sig chi4Sig\n",
    ]
    for res, nchi in LIB_RESIDUES.items():
        nrot = 3 ** nchi
        # rotamer index tuples r1..r4 (1-based, 0 for unused chi)
        idx = np.stack(np.meshgrid(*[np.arange(3)] * nchi, indexing="ij"), -1).reshape(nrot, nchi)
        rcols = np.zeros((nrot, 4), dtype=int)
        rcols[:, :nchi] = idx + 1
        n_bins = 36 * 36
        alpha = np.full(nrot, concentration)
        if well_prior:
            w = _CHI1_PRIOR[idx[:, 0]].copy()
            for k in range(1, nchi):
                w = w * _CHIN_PRIOR[idx[:, k]]
            alpha = concentration * nrot * w / w.sum()
        probs = rng.dirichlet(alpha, size=n_bins)
        centres = np.zeros((n_bins, nrot, 4))
        centres[:, :, :nchi] = _WELLS[idx][None] + rng.normal(0.0, 4.0, size=(n_bins, nrot, nchi))
        centres = (centres + 180.0) % 360.0 - 180.0
        sigmas = np.zeros((n_bins, nrot, 4))
        sigmas[:, :, :nchi] = rng.uniform(6.0, 14.0, size=(n_bins, nrot, nchi))
        lines = []
        for ip, phi in enumerate(_GRID):
            for js, psi in enumerate(_GRID):
                b = (ip % 36) * 36 + (js % 36)
                order = np.argsort(-probs[b], kind="stable")
                for r in order:
                    lines.append(
                        "%s %4d %4d %5d %d %d %d %d %8.6f %6.1f %6.1f %6.1f %6.1f %5.1f %5.1f %5.1f %5.1f\n"
                        % ((res, phi, psi, 100) + tuple(rcols[r]) + (probs[b, r],)
                           + tuple(centres[b, r]) + tuple(sigmas[b, r])))
        chunks.append("".join(lines))
    tmp = str(path) + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        f.writelines(chunks)
    os.replace(tmp, path)
    return path


LIB_VERSION = 2      # bump when the generator changes (cache file name)


def synthetic_library_path(seed=0, cache_dir=None):
    """Return the path of the cached synthetic library for ``seed`` (generated on first use)."""
    cache_dir = Path(cache_dir or os.environ.get("MCSCE_B200_CACHE",
                                                 Path(tempfile.gettempdir()) / "mcsce_b200_cache"))
    d = cache_dir / "SimpleOpt1-5"
    d.mkdir(parents=True, exist_ok=True)
    p = d / ("ALL.bbdep.rotamers.v%d.seed%d.lib" % (LIB_VERSION, seed))
    if not p.exists():
        write_synthetic_bbdep_library(p, seed=seed)
    return p


# ----------------------------------------------------------------------------
# synthetic backbones
# ----------------------------------------------------------------------------

#: axis-to-axis distance of neighbouring helices in the synthetic bundles (Angstrom)
BUNDLE_SPACING = 16.0

STANDARD_20 = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIP", "ILE",
               "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
PTM_SET = ["SEP", "TPO", "PTR", "H2D", "H2E", "M3L", "KCX", "HYP"]

_B_N_CA, _B_CA_C, _B_C_N, _B_C_O, _B_N_H = 1.458, 1.525, 1.329, 1.231, 1.010
_A_N_CA_C, _A_CA_C_N, _A_C_N_CA = np.radians(111.0), np.radians(116.6), np.radians(121.7)
_A_CA_C_O, _A_C_N_H, _A_CA_N_H = np.radians(120.5), np.radians(119.0), np.radians(109.5)


def _place(a, b, c, bond, angle, torsion):
    """NeRF: point d with |cd|=bond, angle(b,c,d)=angle, dihedral(a,b,c,d)=torsion."""
    bc = c - b
    bc /= np.linalg.norm(bc)
    n = np.cross(b - a, bc)
    n /= np.linalg.norm(n)
    m = np.cross(n, bc)
    d2 = np.array([-bond * np.cos(angle),
                   bond * np.sin(angle) * np.cos(torsion),
                   bond * np.sin(angle) * np.sin(torsion)])
    return c + d2[0] * bc + d2[1] * m + d2[2] * n


def build_backbone_segment(seq, phi, psi, omega=180.0):
    """Ideal-geometry backbone for one chain.

    Returns a list of ``(atom_name, residue_index, xyz)`` in the file order the
    reference expects: N CA C O, then H (not PRO/HYP), N-terminal H1 H2 [H3],
    C-terminal OXT.
    """
    L = len(seq)
    phi = np.radians(np.broadcast_to(phi, (L,)).astype(float))
    psi = np.radians(np.broadcast_to(psi, (L,)).astype(float))
    omg = np.radians(float(omega))
    N = np.zeros((L, 3)); CA = np.zeros((L, 3)); C = np.zeros((L, 3))
    N[0] = [0.0, 0.0, 0.0]
    CA[0] = [_B_N_CA, 0.0, 0.0]
    C[0] = CA[0] + _B_CA_C * np.array([-np.cos(_A_N_CA_C), np.sin(_A_N_CA_C), 0.0])
    for i in range(1, L):
        N[i] = _place(N[i - 1], CA[i - 1], C[i - 1], _B_C_N, _A_CA_C_N, psi[i - 1])
        CA[i] = _place(CA[i - 1], C[i - 1], N[i], _B_N_CA, _A_C_N_CA, omg)
        C[i] = _place(C[i - 1], N[i], CA[i], _B_CA_C, _A_N_CA_C, phi[i])
    atoms = []
    for i, res in enumerate(seq):
        atoms.append(("N", i, N[i])); atoms.append(("CA", i, CA[i])); atoms.append(("C", i, C[i]))
        O = _place(N[i], CA[i], C[i], _B_C_O, _A_CA_C_O, psi[i] + np.pi)
        atoms.append(("O", i, O))
        no_h = res in ("PRO", "HYP")
        if i == 0:
            tors = [np.radians(60.0), np.radians(180.0), np.radians(-60.0)]
            names = ["H1", "H2"] if no_h else ["H1", "H2", "H3"]
            for nm, t in zip(names, tors):
                atoms.append((nm, i, _place(C[i], CA[i], N[i], _B_N_H, _A_CA_N_H, t)))
        elif not no_h:
            atoms.append(("H", i, _place(CA[i - 1], C[i - 1], N[i], _B_N_H, _A_C_N_H, 0.0)))
        if i == L - 1:
            atoms.append(("OXT", i, _place(N[i], CA[i], C[i], 1.25, np.radians(117.0), psi[i])))
    return atoms


def _helix_frame(xyz):
    """Rigid transform putting the principal axis of ``xyz`` on +z through the origin."""
    cen = xyz.mean(0)
    u, s, vt = np.linalg.svd(xyz - cen)
    ax = vt[0]
    if np.dot(xyz[-1] - xyz[0], ax) < 0:
        ax = -ax
    z = ax
    x = np.cross([0.0, 1.0, 0.0], z)
    if np.linalg.norm(x) < 1e-3:
        x = np.cross([1.0, 0.0, 0.0], z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return cen, R


def format_pdb(records):
    """records: iterable of (atom_name, res_name, chain, res_num, xyz) -> PDB text."""
    out = []
    for serial, (name, resn, chain, resnum, xyz) in enumerate(records, start=1):
        aname = (" %-3s" % name) if len(name) < 4 else name
        out.append("ATOM  %5d %4s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  \n"
                   % (serial % 100000, aname, resn, chain, resnum, xyz[0], xyz[1], xyz[2], 1.0, 0.0, name[0]))
    out.append("END\n")
    return "".join(out)


def random_sequence(n, seed, alphabet=STANDARD_20, chain_starts=(0,)):
    """Uniform random sequence.  PRO/HYP at a chain's N-terminus is replaced by ALA: the reference
    cannot parameterise it (it names the N-terminal protons H1/H2, libstructure.py:811-816, while the
    force field's NPRO residue carries H2/H3 -> KeyError in libparse.py:447)."""
    rng = np.random.default_rng(seed)
    seq = [alphabet[k] for k in rng.integers(0, len(alphabet), size=n)]
    for p in chain_starts:
        if seq[p] in ("PRO", "HYP"):
            seq[p] = "ALA"
    return seq


def helix_bundle_pdb(seqs, spacing=11.0, seed=0, chain_ids=None, one_chain=False,
                     phi=-57.0, psi=-47.0, cols=3):
    """Backbone-only PDB text for a bundle of ideal alpha helices.

    Each entry of ``seqs`` becomes one helix built with ideal geometry, aligned
    with z, placed on a triangular lattice with ``spacing`` Angstrom between
    neighbouring axes, neighbours antiparallel and each helix spun about its axis
    by a seeded random angle.  Residue numbers are contiguous and globally unique
    (the reference keys its bonded rules on residue number only,
    ``libenergy.py:262-270,337-343``).  With ``one_chain`` every helix carries
    chain id 'A' and only the first/last residue get terminal atoms — the
    connecting loops are omitted, which the reference does not notice: it never
    checks backbone connectivity beyond residue numbering.
    """
    rng = np.random.default_rng(seed)
    records = []
    resnum = 1
    nh = len(seqs)
    for h, seq in enumerate(seqs):
        atoms = build_backbone_segment(seq, phi, psi)
        if one_chain:
            # strip terminal atoms that are interior to the single chain
            keep = []
            for nm, i, x in atoms:
                if nm in ("H1", "H2", "H3") and h != 0:
                    continue
                if nm == "OXT" and h != nh - 1:
                    continue
                keep.append((nm, i, x))
            atoms = keep
            if h != 0 and seq[0] not in ("PRO", "HYP"):
                # interior first residue needs an amide H: put it trans to CA-C about N
                N0 = [x for nm, i, x in atoms if nm == "N" and i == 0][0]
                CA0 = [x for nm, i, x in atoms if nm == "CA" and i == 0][0]
                C0 = [x for nm, i, x in atoms if nm == "C" and i == 0][0]
                Hx = _place(C0, CA0, N0, _B_N_H, _A_C_N_H, np.radians(180.0))
                pos = [k for k, (nm, i, x) in enumerate(atoms) if i == 0 and nm == "O"][0] + 1
                atoms.insert(pos, ("H", 0, Hx))
        xyz = np.array([x for _, _, x in atoms])
        ca = np.array([x for nm, _, x in atoms if nm == "CA"])
        cen, R = _helix_frame(ca)
        xyz = (xyz - cen) @ R.T
        spin = rng.uniform(0, 2 * np.pi)
        cs, sn = np.cos(spin), np.sin(spin)
        Rz = np.array([[cs, -sn, 0], [sn, cs, 0], [0, 0, 1.0]])
        xyz = xyz @ Rz.T
        row, col = divmod(h, cols)
        if (row + col) % 2 == 1:  # antiparallel neighbour: rotate pi about x
            xyz = xyz * np.array([1.0, -1.0, -1.0])
        shift = np.array([col * spacing + (row % 2) * spacing / 2.0, row * spacing * np.sqrt(3) / 2.0, 0.0])
        xyz = xyz + shift
        chain = "A" if one_chain else (chain_ids[h] if chain_ids else "ABCDEFGHIJKLMNOPQRSTUVWXYZ"[h % 26])
        for (nm, i, _), x in zip(atoms, xyz):
            records.append((nm, seq[i], chain, resnum + i, x))
        resnum += len(seq)
    return format_pdb(records)


def config_backbone(cfg, spacing=None):
    """PDB text + description for the BASELINE.json configs (1-5) and small test cases.

    cfg 1/2: 76 residues, one helix-like chain, seed 1.      cfg 3: 300 residues = 6 x 50
    helix bundle, one chain, seed 3.  cfg 4: 150 residues (3 x 50) with >=3 each of the
    PTM set at interior positions, seed 4.  cfg 5: 2000 residues = 8 chains x 250, seed 5.
    """
    if cfg in (1, 2, "76"):
        seq = random_sequence(76, 1)
        return helix_bundle_pdb([seq], seed=1, one_chain=True)
    if cfg in (3, "300"):
        seq = random_sequence(300, 3)
        return helix_bundle_pdb([seq[i:i + 50] for i in range(0, 300, 50)], seed=3, one_chain=True,
                                spacing=spacing or BUNDLE_SPACING)
    if cfg in (4, "ptm150"):
        seq = random_sequence(150, 4)
        rng = np.random.default_rng(40)
        # HYP is a proline: inside an ideal helix its ring always hits the carbonyl O of residue i-4 / i-3 (every trial
        # clashes, in the reference too), so it sits in the first turn of a helix; PRO is kept out of the interior likewise
        first_turn = [1, 2, 51, 52, 101, 102]
        n_each = 3
        pos = [p for p in rng.permutation(np.arange(3, 148)) if p % 50 > 2][: n_each * (len(PTM_SET) - 1)]
        others = [r for r in PTM_SET if r != "HYP"]
        for k, p in enumerate(pos):
            seq[p] = others[k % len(others)]
        for p in rng.permutation(first_turn)[:n_each]:
            seq[p] = "HYP"
        for p in range(150):
            if seq[p] == "PRO" and p % 50 > 2:
                seq[p] = "ALA"
        return helix_bundle_pdb([seq[i:i + 50] for i in range(0, 150, 50)], seed=4, one_chain=True,
                                spacing=spacing or BUNDLE_SPACING)
    if cfg in (5, "2000"):
        seq = random_sequence(2000, 5, chain_starts=range(0, 2000, 250))
        # 8 chains x 250 residues; each chain = 5 helices of 50 (one chain id per 250 residues)
        helices = [seq[i:i + 50] for i in range(0, 2000, 50)]
        return _multichain_bundle(helices, per_chain=5, seed=5, spacing=spacing or BUNDLE_SPACING)
    raise ValueError("unknown config %r" % (cfg,))


def _multichain_bundle(helices, per_chain, seed, spacing=11.0, cols=8):
    """Helix bundle where every ``per_chain`` consecutive helices share a chain id."""
    rng = np.random.default_rng(seed)
    records = []
    resnum = 1
    nh = len(helices)
    for h, seq in enumerate(helices):
        first_in_chain = (h % per_chain == 0)
        last_in_chain = (h % per_chain == per_chain - 1) or h == nh - 1
        atoms = build_backbone_segment(seq, -57.0, -47.0)
        keep = []
        for nm, i, x in atoms:
            if nm in ("H1", "H2", "H3") and not first_in_chain:
                continue
            if nm == "OXT" and not last_in_chain:
                continue
            keep.append((nm, i, x))
        atoms = keep
        if not first_in_chain and seq[0] not in ("PRO", "HYP"):
            N0 = [x for nm, i, x in atoms if nm == "N" and i == 0][0]
            CA0 = [x for nm, i, x in atoms if nm == "CA" and i == 0][0]
            C0 = [x for nm, i, x in atoms if nm == "C" and i == 0][0]
            Hx = _place(C0, CA0, N0, _B_N_H, _A_C_N_H, np.radians(180.0))
            pos = [k for k, (nm, i, x) in enumerate(atoms) if i == 0 and nm == "O"][0] + 1
            atoms.insert(pos, ("H", 0, Hx))
        xyz = np.array([x for _, _, x in atoms])
        ca = np.array([x for nm, _, x in atoms if nm == "CA"])
        cen, R = _helix_frame(ca)
        xyz = (xyz - cen) @ R.T
        spin = rng.uniform(0, 2 * np.pi)
        cs, sn = np.cos(spin), np.sin(spin)
        xyz = xyz @ np.array([[cs, -sn, 0], [sn, cs, 0], [0, 0, 1.0]]).T
        row, col = divmod(h, cols)
        if (row + col) % 2 == 1:
            xyz = xyz * np.array([1.0, -1.0, -1.0])
        xyz = xyz + np.array([col * spacing + (row % 2) * spacing / 2.0, row * spacing * np.sqrt(3) / 2.0, 0.0])
        chain = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"[(h // per_chain) % 26]
        for (nm, i, _), x in zip(atoms, xyz):
            records.append((nm, seq[i], chain, resnum + i, x))
        resnum += len(seq)
    return format_pdb(records)


def small_peptide_pdb(seq, seed=0, phi=-57.0, psi=-47.0):
    """One ideal helix for the given 3-letter sequence (test-sized cases)."""
    return helix_bundle_pdb([list(seq)], seed=seed, one_chain=True, phi=phi, psi=psi)


def text_digest(text):
    return hashlib.sha256(text.encode()).hexdigest()[:16]
