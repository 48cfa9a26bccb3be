"""Minimal structure container for the growth path (host side, stays Python).

Mirrors the parts of the reference's ``Structure`` the growth path and the CLI
touch (``mcsce/libs/libstructure.py:39-570``): PDB / FASTA construction,
``coords`` (float32 copy out, 3-decimal rounding in), terminal-residue flags,
backbone completeness check, side-chain stripping, ``add_side_chain`` atom
bookkeeping, residue-number reordering and PDB output in the reference's line
format (``mcsce/libs/libpdb.py:178-195``).  Storage is column arrays, not the
reference's ``(N,16)`` string table.
"""
from __future__ import annotations

import copy
import os
import re
from pathlib import Path

import numpy as np

from mcsce_b200.chem.tables import BACKBONE_ATOMS

_AA1TO3 = {
    "A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "E": "GLU", "Q": "GLN",
    "G": "GLY", "H": "HIS", "p": "HIP", "e": "HIE", "d": "HID", "I": "ILE", "L": "LEU",
    "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP",
    "Y": "TYR", "V": "VAL",
}
AA3TO1 = {v: k for k, v in _AA1TO3.items()}

_RE_MODEL = re.compile(r"(\n|^)MODEL\s*\d+\s*\n")
_RE_ENDMDL = re.compile(r"\nENDMDL\s*\n")

_COLUMNS = ("record", "serial", "name", "altloc", "resname", "chain", "resseq", "icode",
            "occ", "temp", "segid", "element", "model")


def _round3_f32(coords):
    """What ``Structure.coords = x`` followed by ``.coords`` yields in the reference:
    round to 3 decimals, store as <=8-char text, read back as float32
    (libstructure.py:149-160)."""
    r = np.round(np.asarray(coords), decimals=3)
    if r.size and (np.abs(r).max() >= 1000.0 or r.dtype != np.float32):
        return r.astype("<U8").astype(np.float32)      # the literal route (8-character cells truncate big values)
    return r.astype(np.float32)                         # same values without the text round trip


class Structure:
    """Column-array atom table."""

    def __init__(self, data=None, fasta=None):
        if data is None and fasta is None:
            raise RuntimeError("Please specify either the data source or the FASTA sequence of the structure!")
        self._source = data
        self._fasta = fasta
        self._built = False
        self.cols = {}
        self.xyz = np.zeros((0, 3), dtype=np.float32)

    # -- construction -------------------------------------------------------
    def build(self):
        if self._built:
            return self
        if self._source is not None:
            text = self._source
            if isinstance(text, (str, os.PathLike)) and "\n" not in str(text) and Path(str(text)).exists():
                text = Path(str(text)).read_text()
            elif isinstance(text, bytes):
                text = text.decode()
            self._from_pdb_text(str(text))
        else:
            self._from_fasta(self._fasta)
        self._built = True
        return self

    def _alloc(self, n):
        self.cols = {
            "record": np.full(n, "ATOM", dtype="<U8"), "serial": np.arange(1, n + 1),
            "name": np.full(n, "", dtype="<U8"), "altloc": np.full(n, "", dtype="<U8"),
            "resname": np.full(n, "", dtype="<U8"), "chain": np.full(n, "A", dtype="<U8"),
            "resseq": np.zeros(n, dtype=np.int64), "icode": np.full(n, "", dtype="<U8"),
            "occ": np.full(n, "1.00", dtype="<U8"), "temp": np.full(n, "0.00", dtype="<U8"),
            "segid": np.full(n, "", dtype="<U8"), "element": np.full(n, "", dtype="<U8"),
            "model": np.full(n, "", dtype="<U8"),
        }
        self.xyz = np.zeros((n, 3), dtype=np.float32)

    def _from_pdb_text(self, text):
        m0, m1 = _RE_MODEL.search(text), _RE_ENDMDL.search(text)
        if bool(m0) + bool(m1) == 1:
            raise ValueError("Found MODEL and not ENDMDL, or vice-versa")
        if m0 and m1:
            text = text[m0.span()[1]: m1.span()[0] - 1]
        lines = [ln for ln in text.split("\n") if ln.startswith(("ATOM", "HETATM"))]
        self._alloc(len(lines))
        c = self.cols
        for i, ln in enumerate(lines):
            ln = ln.ljust(80)
            c["record"][i] = ln[0:6].strip()
            c["serial"][i] = int(ln[6:11]) if ln[6:11].strip() else i + 1
            name = ln[12:16].strip()
            if name and name[0].isdigit():
                name = name[1:] + name[0]
            c["name"][i] = name
            c["altloc"][i] = ln[16:17].strip()
            c["resname"][i] = ln[17:20].strip()
            c["chain"][i] = ln[21:22].strip()
            c["resseq"][i] = int(ln[22:26])
            c["icode"][i] = ln[26:27].strip()
            self.xyz[i] = [np.float32(ln[30:38].strip()), np.float32(ln[38:46].strip()), np.float32(ln[46:54].strip())]
            c["occ"][i] = ln[54:60].strip()
            c["temp"][i] = ln[60:66].strip()
            c["segid"][i] = ln[72:76].strip()
            c["element"][i] = ln[76:78].strip()
            c["model"][i] = ln[78:80].strip()

    def _from_fasta(self, seq):
        """Backbone-only table from a 1-letter sequence, zero coordinates
        (libstructure.py:616-629, 804-837): H -> HIP, no amide H on PRO, H1-3 / OXT on the ends."""
        res3 = ["HIP" if r == "H" else _AA1TO3[r] for r in seq]
        rows = []
        for i, res in enumerate(res3):
            if i == 0:
                names = ["N", "CA", "C", "O", "H1", "H2"] + ([] if res == "PRO" else ["H3"])
            else:
                names = ["N", "CA", "C", "O"] + ([] if res == "PRO" else ["H"])
                if i == len(res3) - 1:
                    names.append("OXT")
            rows.extend((nm, res, i + 1) for nm in names)
        self._alloc(len(rows))
        for k, (nm, res, num) in enumerate(rows):
            self.cols["name"][k] = nm
            self.cols["resname"][k] = res
            self.cols["resseq"][k] = num
            self.cols["element"][k] = nm[0]

    # -- views --------------------------------------------------------------
    def __len__(self):
        return self.xyz.shape[0]

    @property
    def coords(self):
        return self.xyz.astype(np.float32, copy=True)

    @coords.setter
    def coords(self, value):
        value = np.asarray(value)
        assert value.shape == self.xyz.shape, (value.shape, self.xyz.shape)
        self.xyz = _round3_f32(value)

    @property
    def atom_labels(self):
        return self.cols["name"].copy()

    @property
    def res_nums(self):
        return self.cols["resseq"].astype(int)

    @property
    def res_labels(self):
        return self.cols["resname"].copy()

    @res_labels.setter
    def res_labels(self, labels):
        self.cols["resname"][:] = labels

    def _chain_residues(self):
        chains = {}
        for ch, num, res in zip(self.cols["chain"], self.cols["resseq"], self.cols["resname"]):
            chains.setdefault(str(ch), {}).setdefault(int(num), str(res))
        return chains

    @property
    def residue_types(self):
        out = []
        for residues in self._chain_residues().values():
            out.extend(residues.values())
        return out

    @property
    def residue_chains(self):
        out = []
        for ch, residues in self._chain_residues().items():
            out.extend([ch] * len(residues))
        return out

    @property
    def residues(self):
        return [int(i) for i in dict.fromkeys(self.cols["resseq"].tolist())]

    @property
    def fasta(self):
        return {ch: "".join(AA3TO1.get(r, "X") for r in residues.values())
                for ch, residues in self._chain_residues().items()}

    def get_terminal_res_atom_arr(self):
        """Boolean per-atom flags for atoms of each chain's first / last residue
        (libstructure.py:163-183)."""
        n_flag = np.zeros(len(self), dtype=bool)
        c_flag = np.zeros(len(self), dtype=bool)
        ch, num = self.cols["chain"], self.cols["resseq"]
        for chain, residues in self._chain_residues().items():
            keys = list(residues.keys())
            n_flag |= (ch == chain) & (num == keys[0])
            c_flag |= (ch == chain) & (num == keys[-1])
        return n_flag, c_flag

    def get_sorted_minimal_backbone_coords(self, mask=None):
        """(3L,3) float32 N,CA,C interleaved (libstructure.py:399-451)."""
        nm = self.cols["name"]
        xyz = self.xyz
        if mask is not None:
            nm, xyz = nm[mask], xyz[mask]
        n, ca, c = xyz[nm == "N"], xyz[nm == "CA"], xyz[nm == "C"]
        assert len(n) == len(ca) == len(c), "backbone N/CA/C counts differ"
        out = np.zeros((3 * len(n), 3), dtype=np.float32)
        out[0::3], out[1::3], out[2::3] = n, ca, c
        return out

    def check_backbone_atom_completeness(self):
        """List of (residue number, atom name) missing from the expected backbone set
        (libstructure.py:453-481)."""
        n_flag, c_flag = self.get_terminal_res_atom_arr()
        per_res = {}
        for nm, num, res, isn, isc in zip(self.cols["name"], self.cols["resseq"], self.cols["resname"], n_flag, c_flag):
            d = per_res.setdefault(int(num), {"label": str(res), "atoms": [], "n": bool(isn), "c": bool(isc)})
            d["atoms"].append(str(nm))
        missing = []
        for num, d in per_res.items():
            no_h = d["label"] in ("PRO", "HYP")
            if d["n"]:
                want = ["N", "CA", "C", "O", "H1", "H2"] + ([] if no_h else ["H3"])
            else:
                want = ["N", "CA", "C", "O"] + ([] if no_h else ["H"]) + (["OXT"] if d["c"] else [])
            missing.extend((num, a) for a in want if a not in d["atoms"])
        return missing

    def _take(self, index):
        out = copy.copy(self)
        out.cols = {k: v[index].copy() for k, v in self.cols.items()}
        out.xyz = self.xyz[index].copy()
        return out

    def copy(self):
        return self._take(np.arange(len(self)))

    def __deepcopy__(self, memo):
        return self.copy()

    def remove_side_chains(self, retain_idxs=()):
        """Copy with only backbone atoms (and whole residues listed in ``retain_idxs``);
        drops the spurious amide H of PRO (libstructure.py:484-496)."""
        keep = np.isin(self.cols["name"], BACKBONE_ATOMS)
        for r in retain_idxs:
            keep |= self.cols["resseq"] == int(r)
        keep &= ~((self.cols["name"] == "H") & (self.cols["resname"] == "PRO"))
        return self._take(np.where(keep)[0])

    def append_atoms(self, names, resname, chain, resseq, xyz, record="ATOM", segid="", elements=None):
        """Append one residue's side-chain block at the end of the table (what
        ``add_side_chain`` does to the atom order, libstructure.py:498-514)."""
        n = len(names)
        add = {
            "record": np.full(n, record, dtype="<U8"), "serial": np.arange(len(self) + 1, len(self) + n + 1),
            "name": np.array(names, dtype="<U8"), "altloc": np.full(n, "", dtype="<U8"),
            "resname": np.full(n, resname, dtype="<U8"), "chain": np.full(n, chain, dtype="<U8"),
            "resseq": np.full(n, int(resseq), dtype=np.int64), "icode": np.full(n, "", dtype="<U8"),
            "occ": np.full(n, "0.00", dtype="<U8"), "temp": np.full(n, "0.00", dtype="<U8"),
            "segid": np.full(n, segid, dtype="<U8"),
            "element": np.array(elements if elements is not None else [a[0] for a in names], dtype="<U8"),
            "model": np.full(n, "", dtype="<U8"),
        }
        for k in self.cols:
            self.cols[k] = np.concatenate([self.cols[k], add[k]])
        self.xyz = np.concatenate([self.xyz, _round3_f32(xyz)])

    def reorder(self, index):
        index = np.asarray(index)
        for k in self.cols:
            self.cols[k] = self.cols[k][index]
        self.xyz = self.xyz[index]
        self.cols["serial"] = np.arange(1, len(index) + 1)

    def reorder_with_resnum(self):
        """Stable sort by residue number (libstructure.py:552-555)."""
        self.reorder(np.argsort(self.res_nums, kind="mergesort"))

    # -- output -------------------------------------------------------------
    def get_PDB(self):
        c = self.cols
        lines = []
        for i in range(len(self)):
            name, elem = str(c["name"][i]).strip(), str(c["element"][i]).strip()
            if len(elem) == 1 and len(name) < 4:
                aname = " {:<3s}".format(name)
            elif len(elem) in (1, 2):
                aname = "{:<4s}".format(name)
            else:
                raise KeyError("Could not format this atom:type -> %s:%s" % (name, elem))
            chain = str(c["chain"][i]).strip()[:1]
            x, y, z = (float(v) for v in self.xyz[i])
            lines.append(
                "{:6s}{:5d} {}{:1s}{:3s} {:1s}{:4d}{:1s}   {:8.3f}{:8.3f}{:8.3f}{:6.2f}{:6.2f}      {:<4s}{:>2s}{:2s}".format(
                    str(c["record"][i]), i + 1, aname, str(c["altloc"][i]), str(c["resname"][i]), chain,
                    int(c["resseq"][i]), str(c["icode"][i]), x, y, z, float(c["occ"][i] or 0.0),
                    float(c["temp"][i] or 0.0), str(c["segid"][i]), elem, str(c["model"][i])))
        return lines

    def pdb_template(self):
        """(template bytes, line length, coordinate offset) for ``mcsce_write_pdbs``: the PDB lines of this
        table with blank coordinate fields, every line padded to 80 characters + newline."""
        saved = self.xyz
        self.xyz = np.zeros_like(saved)
        try:
            lines = self.get_PDB()
        finally:
            self.xyz = saved
        out = []
        for ln in lines:
            ln = ln.ljust(80)[:80]
            out.append(ln[:30] + " " * 24 + ln[54:] + "\n")
        return "".join(out).encode("ascii"), 81, 30

    def write_PDB(self, filename):
        lines = self.get_PDB()
        if not lines:
            raise ValueError("empty structure, nothing to write: %s" % filename)
        with open(filename, "w") as fh:
            fh.write("\n".join(lines))
            fh.write("\n")


def read_structure_and_check(file_name, retain_idx=(), verbose=False):
    """Read a PDB, rename HIS -> HIP, report missing backbone atoms, strip side chains
    (mcsce/cli.py:40-69)."""
    s = Structure(file_name).build()
    if (s.cols["resname"] == "HIS").any():
        s.cols["resname"][s.cols["resname"] == "HIS"] = "HIP"
        print("HIS residues modified to HIP\n")
    missing = s.check_backbone_atom_completeness()
    if missing:
        if verbose:
            print("WARNING! These atoms are missing from the current backbone structure [%s]:" % file_name
                  + "".join("\n%s %s" % m for m in missing) + "\n")
        else:
            print("WARNING! %d missing atoms in total" % len(missing))
    return s.remove_side_chains(retain_idx)
