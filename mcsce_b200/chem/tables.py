"""Static chemistry tables for the growth path (force-field parameters, residue
templates, bonded topology, chi-rotation definitions, PTM rotamers).

The reference derives these at import time from its data files
(``mcsce/core/build_definitions.py:67-166`` XML -> dict, ``:169-312`` template
topology, ``:362-456`` bonds-apart expansion, ``:775-788`` templates;
``mcsce/core/rotamer_library.py:180-241`` PTM library) and re-derives per-pair
masks for every residue (``mcsce/libs/libenergy.py:80-106``).  Here the same
information is parsed once by :func:`build_tables_from_reference_data` (own
parsers, no reference import) into one packed file,
``mcsce_b200/data/chem_tables.json.gz``, which is what the package loads at run
time - the reference tree does not exist on the GPU box.  In a deployment the
packed file is regenerated from the ``mcsce`` package's own data directory.

Everything downstream (host tables for the CUDA kernels, the oracle) consumes
:class:`ChemTables`.
"""
from __future__ import annotations

import gzip
import json
import re
import xml.etree.ElementTree as ET
from collections import deque
from functools import lru_cache
from pathlib import Path

import numpy as np

#: atoms that ``remove_side_chains`` keeps and that are never "side chain"
#: (mcsce/core/definitions.py:17).  HA/HA2/HA3 are *not* in this set, so they are
#: grown with the side chain.
BACKBONE_ATOMS = ("N", "C", "CA", "O", "OXT", "H", "H1", "H2", "H3")

#: Tsai 1999 radii keyed by the first character of the atom name
#: (mcsce/core/definitions.py:76-83, used by libenergy.py:526-537).
VDW_RADII = {"C": 1.7, "H": 1.0, "N": 1.625, "O": 1.480, "P": 1.871, "S": 1.782}
CLASH_ELEMENTS = ("C", "H", "N", "O", "P", "S")
#: clash distance = CLASH_FACTOR * (r_i + r_j)  (libenergy.py:152)
CLASH_FACTOR = 0.6
#: extra scale on 1-4 Lennard-Jones on top of lj14scale (libenergy.py:171-174)
LJ14_EXTRA_SCALE = 0.2
#: Coulomb constant kJ*A/mol/e^2 and dielectric factor (libenergy.py:521-522)
COULOMB_K = 1389.35
DIELECTRIC_SCALE = 0.25

#: PTM residue -> parent residue whose bbdep rotamers are used (definitions.py:33-37)
PTM_PARENT = {
    "AGM": "ARG", "MEN": "ASN", "BHD": "ASP", "SMC": "CYS", "MGN": "GLN", "CGU": "GLU",
    "H2E": "HIS", "H2D": "HIS", "H1E": "HIS", "H1D": "HIS", "HYP": "PRO", "SEP": "SER",
    "TPO": "THR", "S1P": "SER", "T1P": "THR", "Y1P": "TYR", "PTR": "TYR", "TYS": "TYR",
    "LYZ": "LYS", "ALY": "LYS", "M3L": "LYS", "KCX": "LYS",
}
#: protonation variants that share another residue's PTM rotamers (definitions.py:38)
PTM_ROTAMER_ALIAS = {"H1E": "H2D", "H1D": "H2D", "H2E": "H2D", "S1P": "SEP", "T1P": "TPO", "Y1P": "PTR"}

#: inter-residue bonded rules, residue n -> residue n+1, keyed by atom name in n
#: (build_definitions.py:460-482).
INTER_BOND_1 = {"C": ("N",)}
INTER_BOND_LE2 = {"C": ("CA", "H", "N"), "CA": ("N",), "O": ("N",)}
INTER_BOND_EQ3 = {
    "N": ("N",), "CA": ("H", "CA"), "C": ("HA", "HA2", "HA3", "CB", "C", "CH3"),
    "O": ("CA", "H"), "HA": ("N",), "HA2": ("N",), "HA3": ("N",), "CB": ("N",),
}

# ---------------------------------------------------------------------------
# chi-rotation rules (restated from libscbuild.py:87-206 as data)
# ---------------------------------------------------------------------------
# Template atoms are addressed by *position* in the template file, not by IUPAC
# name: positions 0,1 = N,CA; 4..9 = the chain CB,G,D,E,Z,H positions.
_CHAIN = (0, 1, 4, 5, 6, 7, 8, 9)
#: atoms (by name) that stop moving at each successive generic chi
_GENERIC_DROP = (
    ("HA", "CB2", "HB21", "HB22", "HB23"),
    ("HB", "HB2", "HB3", "CB", "CG2", "HG21", "HG22", "HG23", "CB1", "HB12", "HB13", "OB"),
    ("HG", "HG2", "HG3", "CG", "CG1"),
    ("HD", "HD2", "HD3", "CD", "SD"),
    ("HE", "HE1", "HE2", "HE3", "CD"),
)
_EXTRA_DROP = {  # (chi number, residue) -> more names
    (3, "MEN"): ("OD1",), (3, "CGU"): ("OE21", "OE22", "CD2"),
    (3, "SEP"): ("OG",), (3, "S1P"): ("OG",), (3, "TPO"): ("OG1",), (3, "T1P"): ("OG1",),
    (4, "LYZ"): ("OH", "HH"), (4, "AGM"): ("CE2", "HE21", "HE22", "HE23"),
}
#: residues whose chi3 (and chi4) use explicit position lists:
#: residue -> list of (dihedral positions, moving positions); -1 = last template atom.
_SPECIAL_CHI = {
    "PTR": [((9, 10, 11, 12), (12, 13, 14, 15)), ((10, 11, 12, 13), (13, 14, 15))],
    "Y1P": [((9, 10, 11, 12), (12, 13, 14, 15, -1)), ((10, 11, 12, 13), (13, 14, 15, -1))],
    "H2D": [((5, 7, 10, 11), (11, 12, 13))],
    "H1D": [((5, 7, 10, 11), (11, 12, 13, -1))],
    # H1E's extra proton is *not* moved: the reference tests for 'H1D' in this branch
    # (libscbuild.py:162), reproduced here on purpose.
    "H2E": [((6, 9, 10, 11), (11, 12, 13))],
    "H1E": [((6, 9, 10, 11), (11, 12, 13))],
}
NO_CHI_RESIDUES = ("ALA", "GLY")


def _amber_name(name):
    """Leading-digit hydrogen names are rotated to the end (libstructure.py:795-801)."""
    if name and name[0].isdigit():
        return name[1:] + name[0]
    return name


def _parse_template_pdb(path):
    names, coords, extra = [], [], {"resname": None, "elements": [], "tails": []}
    for line in Path(path).read_text().splitlines():
        if line.startswith(("ATOM", "HETATM")):
            line = line.ljust(80)
            names.append(_amber_name(line[12:16].strip()))
            coords.append([line[30:38].strip(), line[38:46].strip(), line[46:54].strip()])
            extra["resname"] = extra["resname"] or line[17:20].strip()
            extra["elements"].append(line[76:78].strip())
            extra["tails"].append(line[78:80].strip())
    return names, coords, extra


def _parse_forcefield_xml(paths):
    """Atom-type LJ parameters and per-residue (charge, type); later files override."""
    type_params, residues, scales = {}, {}, {}
    for p in paths:
        root = ET.fromstring(Path(p).read_text())
        for child in root:
            if child.tag == "Residues":
                for res in child:
                    table = residues.setdefault(res.attrib["name"], {})
                    for atom in res:
                        if atom.tag == "Atom":
                            table[atom.attrib["name"]] = [atom.attrib["charge"], atom.attrib["type"]]
            elif child.tag == "NonbondedForce":
                scales["coulomb14scale"] = child.attrib["coulomb14scale"]
                scales["lj14scale"] = child.attrib["lj14scale"]
                for atom in child:
                    if atom.tag == "Atom":
                        type_params[atom.attrib["type"]] = [atom.attrib["sigma"], atom.attrib["epsilon"]]
    return type_params, residues, scales


def _parse_ptm_library(path):
    """SIDEpro-format PTM rotamers: info[res] = [n_chi, n_extra, dependence], data[(res, lo)] rows."""
    info, data = {}, {}
    res = None
    for raw in Path(path).read_text().splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("#"):
            m = re.fullmatch(r"# (\w{3})", line)
            if m:
                res = m.group(1)
                info[res] = []
            elif line.startswith("# Number of chi") or line.startswith("# Number of additional chi"):
                info[res].append(int(line.split()[-1]))
            elif line.startswith("# Dependence"):
                info[res].append(int(line.split()[-1]))
            elif line.startswith("# No Dependence"):
                info[res].append(0)
            continue
        if res is not None and line.startswith(res):
            items = line.split()
            if info[res][-1] > 0:
                key, row = "%s|%s" % (res, float(items[1])), [float(v) for v in items[3:]]
            else:
                key, row = "%s|%s" % (res, -360.0), [float(v) for v in items[1:]]
            data.setdefault(key, []).append(row)
    return info, data


def build_tables_from_reference_data(core_dir, out_path=None):
    """Parse ``<core_dir>/data/*.xml``, ``ptm_rotamer.lib`` and ``sidechain_templates/**.pdb``
    and write the packed table file.  ``core_dir`` is the ``mcsce/core`` directory of an
    installed reference."""
    core_dir = Path(core_dir)
    type_params, residues, scales = _parse_forcefield_xml(
        [core_dir / "data" / "protein.ff14SB.xml", core_dir / "data" / "phosaa14SB.xml"])
    templates = {}
    for sub in ("amber_names", "ptm_templates"):
        for pdb in sorted((core_dir / "sidechain_templates" / sub).glob("*.pdb")):
            names, coords, extra = _parse_template_pdb(pdb)
            templates[pdb.stem.upper()] = dict(names=names, coords=coords, **extra)
    info, data = _parse_ptm_library(core_dir / "data" / "ptm_rotamer.lib")
    packed = {
        "format": 1,
        "source": "THGLab/MCSCE core/data + core/sidechain_templates (parsed, not copied)",
        "scales": scales, "atom_types": type_params, "residues": residues,
        "templates": templates, "ptm_info": info, "ptm_data": data,
    }
    out_path = Path(out_path or default_tables_path())
    out_path.parent.mkdir(parents=True, exist_ok=True)
    with gzip.GzipFile(out_path, "wb", mtime=0) as f:
        f.write(json.dumps(packed, sort_keys=True, separators=(",", ":")).encode())
    return out_path


def default_tables_path():
    return Path(__file__).resolve().parents[1] / "data" / "chem_tables.json.gz"


# ---------------------------------------------------------------------------
# run-time view
# ---------------------------------------------------------------------------

def dihedral_f32(p):
    """Dihedral of four float32 points, evaluated in float32 with the same
    formula the reference uses (libcalc.py:271-304: unit normals of the two
    planes, ``-arctan2(sin, cos)``)."""
    p = np.asarray(p, dtype=np.float32)
    crds = p.T
    q = crds[:, 1:] - crds[:, :-1]
    cross = np.cross(q[:, :-1], q[:, 1:], axis=0)
    unitary = cross / np.linalg.norm(cross, axis=0)
    u0 = unitary[:, :-1]
    u1 = unitary[:, 1:]
    u3 = q[:, 1:-1] / np.linalg.norm(q[:, 1:-1], axis=0)
    u2 = np.cross(u3, u1, axis=0)
    cos_t = np.diagonal(np.matmul(u0.T, u1))
    sin_t = np.diagonal(np.matmul(u0.T, u2))
    return -np.arctan2(sin_t, cos_t)


class ResidueTemplate:
    """One residue template: names, float32 coords (CA at the origin), side-chain
    positions and the chi rotation program."""

    def __init__(self, res, names, coords_txt, resname=None, elements=None, tails=None):
        self.res = res
        self.resname = resname or res          # label the grown atoms carry (Y1P's file says TYR)
        self.names = list(names)
        self.elements = list(elements) if elements else [n[0] for n in names]
        self.tails = list(tails) if tails else [""] * len(names)
        self.coords = np.array([[np.float32(v) for v in row] for row in coords_txt], dtype=np.float32)
        self.n_atoms = len(self.names)
        self.sc_idx = np.array([i for i, n in enumerate(self.names) if n not in BACKBONE_ATOMS], dtype=np.int64)
        self.n_sc = len(self.sc_idx)
        self.chis = self._chi_program()
        self.n_sound = self._count_sound()

    def _count_sound(self):
        """Leading rotations whose dihedral is unchanged by every earlier rotation (each of its four atoms
        either moves with that rotation or sits on its axis).  That holds for every real chi angle and makes
        the template dihedral a constant; it fails for the position-based "extra" entries the reference
        would apply if handed more angles than the residue has (e.g. a 4th angle for ASN reaches the
        backbone H), which no rotamer library row ever does.  The CUDA path applies sound rotations only."""
        for k, chi in enumerate(self.chis):
            for prev in self.chis[:k]:
                fixed_ok = set(prev["axis"])
                moved = set(prev["moving"])
                atoms = set(chi["dihedral"])
                if not (atoms <= (moved | fixed_ok)) and (atoms & moved):
                    return k
        return len(self.chis)

    def _chi_program(self):
        """List of dicts {dihedral:(a,b,c,d), axis:(from,to), pivot, moving:[positions], ori:float}.

        ``ori`` is the dihedral of the *unrotated* template measured in float32.  The
        reference re-measures it after every earlier rotation (libscbuild.py:113,167,
        183,197); mathematically that value is invariant (earlier rotations move the
        later dihedral's atoms rigidly about an axis through two of them), so the
        re-measurement only adds float32 noise of ~1e-7 rad.
        """
        if self.res in NO_CHI_RESIDUES or self.res == "PCA":
            return []
        nm = self.names
        last = self.n_atoms - 1
        prog = []
        moving = list(self.sc_idx)
        have = lambda *pos: all(0 <= p < self.n_atoms for p in pos)  # noqa: E731
        special = _SPECIAL_CHI.get(self.res)
        for k in range(5):
            if special is not None and k >= 2:
                j = k - 2
                if j >= len(special):
                    break
                dih, mov = special[j]
                mov = [last if m == -1 else m for m in mov]
                if not have(*dih) or not have(*mov):
                    break
                prog.append(self._entry(dih, (dih[1], dih[2]), mov))
                continue
            dih = _CHAIN[k:k + 4]
            if not have(*dih):
                break
            drop = set(_GENERIC_DROP[k]) | set(_EXTRA_DROP.get((k + 1, self.res), ()))
            moving = [m for m in moving if nm[m] not in drop]
            prog.append(self._entry(dih, (dih[1], dih[2]), list(moving)))
        return prog

    def _entry(self, dih, axis, moving):
        ori = float(dihedral_f32(self.coords[list(dih)])[0])
        return {"dihedral": tuple(int(d) for d in dih), "axis": (int(axis[0]), int(axis[1])),
                "pivot": int(axis[1]), "moving": [int(m) for m in moving], "ori": ori}


class ChemTables:
    """Parsed view of the packed table file."""

    def __init__(self, packed):
        self.lj14scale = float(packed["scales"]["lj14scale"])
        self.coulomb14scale = float(packed["scales"]["coulomb14scale"])
        self.atom_types = {k: (float(v[0]), float(v[1])) for k, v in packed["atom_types"].items()}
        self.residues = {r: {a: (float(c), t) for a, (c, t) in atoms.items()}
                         for r, atoms in packed["residues"].items()}
        self.templates = {r: ResidueTemplate(r, t["names"], t["coords"], t.get("resname"), t.get("elements"), t.get("tails")) for r, t in packed["templates"].items()}
        self.ptm_info = {k: list(v) for k, v in packed["ptm_info"].items()}
        self.ptm_data = {}
        for key, rows in packed["ptm_data"].items():
            res, lo = key.split("|")
            self.ptm_data[(res, float(lo))] = np.array(rows, dtype=np.float64)
        self.bonds = {r: self._template_bonds(t) for r, t in self.templates.items()}
        self._dist_cache = {}

    # -- bonded topology ----------------------------------------------------
    def _template_bonds(self, tmpl):
        """Covalent neighbours per atom name: pairs <= 1.65 A apart in the template plus the
        S/terminal special cases (build_definitions.py:214-273)."""
        x = tmpl.coords.astype(np.float64)
        d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
        nb = {n: [] for n in tmpl.names}
        for i in range(tmpl.n_atoms):
            for j in range(i + 1, tmpl.n_atoms):
                if d[i, j] <= 1.65:
                    nb[tmpl.names[i]].append(tmpl.names[j])
                    nb[tmpl.names[j]].append(tmpl.names[i])

        def link(a, b):
            if b not in nb[a]:
                nb[a].append(b)
            if a not in nb[b]:
                nb[b].append(a)

        if tmpl.res == "CYS":
            link("CB", "SG")
        elif tmpl.res == "MET":
            link("CG", "SD"); link("SD", "CE")
            nb["SD"] = ["CG", "CE"]
        elif tmpl.res == "SMC":
            link("CB", "SG"); link("SG", "CS")
            nb["SG"] = ["CB", "CS"]
        elif tmpl.res == "TYS":
            nb["OH"].append("S"); nb["S"].extend(("HO3", "O3S"))
        # OXT mirrors O; H1..H3 hang off N (add_OXT_to_connectivity / add_Nterm_H_connectivity)
        nb["OXT"] = list(nb["O"])
        for a in nb["OXT"]:
            nb[a].append("OXT")
        nb["N"].extend(("H1", "H2", "H3"))
        for h in ("H1", "H2", "H3"):
            nb[h] = ["N"]
        return nb

    def bond_distance(self, res, a, b):
        """Number of bonds (1, 2, 3) between atoms ``a`` and ``b`` of one residue, or 0 if
        farther apart / unknown.  Reproduces what the reference's three intra-residue masks
        encode after their overrides (libenergy.py:80-106, 171-178): <=2 excluded, ==3 scaled."""
        key = (res, a)
        dist = self._dist_cache.get(key)
        if dist is None:
            nb = self.bonds[res]
            if a not in nb:
                raise KeyError("atom %s not in topology of %s" % (a, res))
            dist = {a: 0}
            dq = deque([a])
            while dq:
                u = dq.popleft()
                if dist[u] == 3:
                    continue
                for v in nb.get(u, ()):
                    if v not in dist:
                        dist[v] = dist[u] + 1
                        dq.append(v)
            self._dist_cache[key] = dist
        if b not in self.bonds[res]:
            raise KeyError("atom %s not in topology of %s" % (b, res))
        return dist.get(b, 0)

    # -- per-atom force-field parameters -------------------------------------
    def atom_params(self, res_label, atom_name, is_nterm, is_cterm):
        """(charge, sigma_nm, epsilon) for one atom; terminal residues use the N*/C* residue
        entries (libparse.py:419-453)."""
        res = ("N" + res_label) if is_nterm else (("C" + res_label) if is_cterm else res_label)
        try:
            charge, atype = self.residues[res][atom_name]
        except KeyError:
            raise KeyError("no force-field entry for atom %s of residue %s" % (atom_name, res))
        sigma, eps = self.atom_types[atype]
        return charge, sigma, eps


@lru_cache(maxsize=2)
def load_tables(path=None):
    path = Path(path or default_tables_path())
    if not path.exists():
        raise FileNotFoundError(
            "%s missing: run `python -m mcsce_b200.chem.tables <mcsce/core dir>` to build it" % path)
    with gzip.open(path, "rb") as f:
        return ChemTables(json.loads(f.read().decode()))


if __name__ == "__main__":
    import sys

    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/mcsce/core"
    print(build_tables_from_reference_data(src))
