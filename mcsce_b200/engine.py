"""GrowthEngine: Python host of the CUDA growth path.

Packs a :class:`mcsce_b200.system.GrowthSystem` into the C-ABI descriptors, owns the device
buffers (torch tensors: device memory and streams only) and calls ``libmcsce_b200.so``.
No compute happens here and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from mcsce_b200 import _lib
from mcsce_b200._lib import (MAX_CHI, MAX_CHI_ROT, MAX_SC, MAX_SP_ENV, MAX_TMPL, BackboneDesc, GrowOpts,
                             SystemDesc, check, load_library)
from mcsce_b200.chem.tables import load_tables
from mcsce_b200.libs.placement import placement_quaternions_batch
from mcsce_b200.system import assign_trials, placement_backbone


def _ptr(a, typ):
    return a.ctypes.data_as(C.POINTER(typ))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class GrowthEngine:
    def __init__(self, device=0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("mcsce_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load_library()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        check(self.lib.mcsce_create(C.byref(h), self.device_index))
        self.h = h
        self.system = None
        self._keep = []
        self._workspace = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.mcsce_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tables ---------------------------------------------------------------
    def set_system(self, sysm, tables=None):
        t = tables or load_tables()
        self.system = sysm
        self.tables = t
        types = sorted({st.resname for st in sysm.steps if st.kind != "skip"})
        self.type_index = {r: i for i, r in enumerate(types)}
        nt = max(1, len(types))
        tmpl_n = np.zeros(nt, np.int32); tmpl_xyz = np.zeros((nt, MAX_TMPL, 3), np.float32)
        tmpl_sc = np.full((nt, MAX_SC), -1, np.int32); tmpl_nchi = np.zeros(nt, np.int32)
        tmpl_axis = np.zeros((nt, MAX_CHI_ROT, 2), np.int32); tmpl_move = np.zeros((nt, MAX_CHI_ROT), np.uint32)
        tmpl_ori = np.zeros((nt, MAX_CHI_ROT), np.float64)
        for r, i in self.type_index.items():
            tm = t.templates[r]
            assert tm.n_atoms <= MAX_TMPL and tm.n_sc <= MAX_SC, r
            tmpl_n[i] = tm.n_atoms
            tmpl_xyz[i, :tm.n_atoms] = tm.coords
            tmpl_sc[i, :tm.n_sc] = tm.sc_idx
            prog = tm.chis[:min(MAX_CHI_ROT, tm.n_sound)]
            tmpl_nchi[i] = len(prog)
            for k, chi in enumerate(prog):
                tmpl_axis[i, k] = chi["axis"]
                tmpl_move[i, k] = sum(1 << m for m in chi["moving"])
                tmpl_ori[i, k] = chi["ori"]
        L = sysm.n_res
        kind = np.array([{"skip": 0, "rigid": 1, "grow": 2}[st.kind] for st in sysm.steps], np.int32)
        stype = np.array([self.type_index.get(st.resname, 0) for st in sysm.steps], np.int32)
        n_sc = np.array([st.n_sc for st in sysm.steps], np.int32)
        sc_start = np.array([st.sc_start for st in sysm.steps], np.int32)
        n_sp_env = np.zeros(L, np.int32); sp_env = np.zeros((L, MAX_SP_ENV), np.int32)
        sp_off = np.zeros(L, np.int32); n_sp = np.zeros(L, np.int32)
        ks, partners, coefs = [], [], []
        off = 0
        for st in sysm.steps:
            if st.kind != "grow":
                continue
            ne = len(st.special_env)
            if ne > MAX_SP_ENV:
                raise ValueError("residue %d has %d special environment atoms (max %d)" % (st.resnum, ne, MAX_SP_ENV))
            n_sp_env[st.index] = ne
            sp_env[st.index, :ne] = st.special_env
            sp_off[st.index] = off; n_sp[st.index] = len(st.sp_k)
            ks.append(st.sp_k); partners.append(st.sp_partner); coefs.append(st.sp_coef)
            off += len(st.sp_k)
        cat = lambda xs, dt, shape: (np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dt))  # noqa: E731
        sp_k = cat(ks, np.int32, (0,)); sp_partner = cat(partners, np.int32, (0,)); sp_coef = cat(coefs, np.float64, (0, 4))
        charge = sysm.charge if "coulomb" in sysm.terms else np.zeros_like(sysm.charge)
        bb_i, bb_j, bb_coef = sysm.bb_pairs
        arrays = dict(
            slot_of_atom=_i32(sysm.slot_of_atom), cls_of_atom=_i32(sysm.cls_of_atom), charge=_f64(charge),
            resnum=_i32(sysm.resnums), cls_start=_i32(sysm.cls_start), cls_active=_i32(sysm.cls_active),
            cls_s2=_f64(sysm.cls_s2), cls_e4=_f64(sysm.cls_e4), cls_thr=_f64(sysm.cls_thr),
            tmpl_n_atoms=tmpl_n, tmpl_xyz=_f32(tmpl_xyz), tmpl_sc_pos=_i32(tmpl_sc), tmpl_n_chi=tmpl_nchi,
            tmpl_axis=_i32(tmpl_axis), tmpl_move=np.ascontiguousarray(tmpl_move, np.uint32), tmpl_ori=_f64(tmpl_ori),
            step_kind=kind, step_type=stype, step_n_sc=n_sc, step_sc_start=sc_start, step_n_sp_env=n_sp_env,
            step_sp_env=_i32(sp_env), step_sp_off=sp_off, step_n_sp=n_sp, sp_k=sp_k, sp_partner=sp_partner,
            sp_coef=_f64(sp_coef), bb_i=_i32(bb_i), bb_j=_i32(bb_j), bb_coef=_f64(bb_coef))
        d = SystemDesc()
        d.n_atoms, d.n_input, d.n_res, d.n_cls = sysm.n_atoms, sysm.n_input, L, sysm.n_cls
        d.n_types, d.n_sp_total, d.n_bb_pairs = len(types), len(sp_k), len(bb_i)
        for name, arr in arrays.items():
            typ = dict(SystemDesc._fields_)[name]._type_
            setattr(d, name, _ptr(arr, typ))
        check(self.lib.mcsce_set_system(self.h, C.byref(d)))
        self._keep = [arrays, d]
        self._step_kind = kind
        self._has_backbone = False
        return self

    def set_backbone(self, backbone_coords):
        """Upload the backbone-dependent tables: trial sets, template placement, coordinates."""
        sysm, t = self.system, self.tables
        backbone_coords = np.asarray(backbone_coords, dtype=np.float32)
        assert backbone_coords.shape == (sysm.n_input, 3), "backbone atom count mismatch"
        L = sysm.n_res
        if sysm.rotamer_library is not None:
            assign_trials(sysm, backbone_coords, t)      # trial sets follow the backbone's phi/psi (side_chain_builder.py:144-147,171)
        bb = placement_backbone(backbone_coords, sysm)
        q1 = np.zeros((L, 4)); q2 = np.zeros((L, 4)); ca = np.zeros((L, 3), np.float32)
        q1[:, 0] = q2[:, 0] = 1.0
        n_tr = np.zeros(L, np.int32); tr_off = np.zeros(L, np.int32)
        n_lib = np.zeros(L, np.int32); n_chi = np.zeros(L, np.int32)
        means, sigmas, probs = [], [], []
        off = 0
        placed = [st for st in sysm.steps
                  if st.kind != "skip" and not (st.kind == "grow" and st.chi_mean is None)]   # (no trial sets: scoring-only use)
        if placed:      # template placement of all residues at once (same arithmetic as the per-residue routine)
            idx = np.array([st.index for st in placed])
            tn = np.stack([t.templates[st.resname].coords[0] for st in placed])
            tc = np.stack([t.templates[st.resname].coords[2] for st in placed])
            q1[idx], q2[idx], ca[idx] = placement_quaternions_batch(bb[idx], tn, tc)
        for st in placed:
            if st.kind != "grow":
                continue
            T = len(st.prob)
            m = np.zeros((T, MAX_CHI)); s = np.zeros((T, MAX_CHI))
            m[:, :st.n_chi_lib] = st.chi_mean; s[:, :st.n_chi_lib] = st.chi_sigma
            means.append(m); sigmas.append(s); probs.append(st.prob)
            n_tr[st.index], tr_off[st.index], n_lib[st.index], n_chi[st.index] = T, off, st.n_chi_lib, st.n_chi
            off += T
        arrays = dict(
            step_n_trials=n_tr, step_trial_off=tr_off, step_n_chi_lib=n_lib, step_n_chi=n_chi,
            chi_mean=_f64(np.concatenate(means) if means else np.zeros((0, MAX_CHI))),
            chi_sigma=_f64(np.concatenate(sigmas) if sigmas else np.zeros((0, MAX_CHI))),
            prob=_f64(np.concatenate(probs) if probs else np.zeros(0)),
            place_q1=_f64(q1), place_q2=_f64(q2), place_ca=_f32(ca), input_xyz=_f32(backbone_coords))
        d = BackboneDesc()
        d.n_trials_total = off
        for name, arr in arrays.items():
            typ = dict(BackboneDesc._fields_)[name]._type_
            setattr(d, name, _ptr(arr, typ))
        check(self.lib.mcsce_set_backbone(self.h, C.byref(d)))
        self.n_trials_total = off
        self.trial_off = tr_off
        self.n_trials = n_tr
        self._has_backbone = True
        return self

    # -- growth -----------------------------------------------------------------
    def _ws(self, nbytes):
        torch = self.torch
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._workspace

    def workspace_bytes(self, n_conf):
        return int(self.lib.mcsce_workspace_bytes(self.h, int(n_conf)))

    def grow(self, n_conf, beta, seed=0, conf_id0=0, noise_deg=0.5, chis=None, uniforms=None, dump=False,
             env_split=0, host_out=False, stream=None, no_dedup=False, no_graph=False, no_screen=False, oris=None,
             out_ptrs=None, hpmf=None):
        """Grow ``n_conf`` conformers.  Returns dict(coords (C,N,3) f32 growth order, energy (C,) f64,
        ok (C,) uint8 [, trial_energy, chis, selected]); torch CUDA tensors, or NumPy arrays (pinned
        staging inside the C call) with ``host_out``.  ``out_ptrs`` = (coords, energy, ok) raw device addresses: the
        results are stored there instead (e.g. this rank's slice of rank 0's peer-mapped buffer, ``sharding.PeerResults``).
        ``hpmf``: a :class:`mcsce_b200.hpmf.HpmfCalculator` for this atom list - adds the hydrophobic PMF of (conformer so far
        + trial) to every trial that passed the clash filter, inside the fused loop (``terms=[..., "hpmf"]``)."""
        torch = self.torch
        assert self._has_backbone, "set_backbone first"
        if any(st.kind == "grow" and st.chi_mean is None for st in self.system.steps):
            raise ValueError("system was built without a rotamer library: no trial sets to grow from")
        C_ = int(n_conf)
        N = self.system.n_atoms
        o = GrowOpts()
        o.n_conf, o.conf_id0, o.seed, o.beta, o.noise_deg, o.env_split = C_, int(conf_id0), int(seed), float(beta), float(noise_deg), int(env_split)
        o.no_dedup = 1 if no_dedup else 0
        o.no_graph = int(no_graph)          # False/0 automatic, True/1 direct per-kernel path, 2 graph replay, 3 persistent cluster kernel
        o.no_screen = 1 if no_screen else 0
        keep = []
        if hpmf is not None:
            assert hpmf.n_atoms == N, "the hpmf calculator was built for another atom list"
            o.hpmf = hpmf.h.value if hasattr(hpmf.h, "value") else int(hpmf.h)
            keep.append(hpmf)
        if chis is not None:
            ch = torch.as_tensor(np.ascontiguousarray(chis, np.float64), device=self.device)
            assert ch.numel() == C_ * self.n_trials_total * MAX_CHI, "chis must be (n_conf, n_trials_total, MAX_CHI)"
            o.d_chis = ch.data_ptr(); keep.append(ch)
        if uniforms is not None:
            un = torch.as_tensor(np.ascontiguousarray(uniforms, np.float64), device=self.device)
            assert un.numel() == C_ * self.system.n_res
            o.d_uniform = un.data_ptr(); keep.append(un)
        if oris is not None:      # parity replay: the dihedrals the reference measured before each rotation (NaN = constant)
            assert chis is not None, "oris replay the measured dihedrals of given chi rows"
            orr = torch.as_tensor(np.ascontiguousarray(oris, np.float64), device=self.device)
            assert orr.numel() == C_ * self.n_trials_total * MAX_CHI_ROT, "oris must be (n_conf, n_trials_total, MAX_CHI_ROT)"
            o.d_ori = orr.data_ptr(); keep.append(orr)
        out = {}
        if dump:
            out["trial_energy"] = torch.full((C_, self.n_trials_total), float("nan"), dtype=torch.float64, device=self.device)
            out["chis"] = torch.zeros((C_, self.n_trials_total, MAX_CHI), dtype=torch.float64, device=self.device)
            out["selected"] = torch.full((C_, self.system.n_res), -2, dtype=torch.int32, device=self.device)
            o.d_trial_energy, o.d_chis_out, o.d_selected = (out[k].data_ptr() for k in ("trial_energy", "chis", "selected"))
        s_ptr = C.c_void_p(stream.cuda_stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream)
        base = self.workspace_bytes(C_)
        if host_out:
            extra = C_ * (N * 12 + 9) + 1024
            ws = self._ws(base + extra)
            key = (C_, N)
            if getattr(self, "_pinned_key", None) != key:       # pinned staging is reused between calls
                self._pinned = (torch.empty((C_, N, 3), dtype=torch.float32).pin_memory(),
                                torch.empty(C_, dtype=torch.float64).pin_memory(),
                                torch.empty(C_, dtype=torch.uint8).pin_memory())
                self._pinned_key = key
            coords, energy, ok = self._pinned
            check(self.lib.mcsce_grow_host(self.h, C.byref(o), C.c_void_p(ws.data_ptr()), ws.numel(),
                                           C.c_void_p(coords.data_ptr()), C.c_void_p(energy.data_ptr()),
                                           C.c_void_p(ok.data_ptr()), s_ptr))
            out.update(coords=coords.numpy(), energy=energy.numpy(), ok=ok.numpy())   # views of the pinned staging: copy to keep
        elif out_ptrs is not None:
            ws = self._ws(base)
            check(self.lib.mcsce_grow(self.h, C.byref(o), C.c_void_p(ws.data_ptr()), ws.numel(),
                                      C.c_void_p(int(out_ptrs[0])), C.c_void_p(int(out_ptrs[1])),
                                      C.c_void_p(int(out_ptrs[2])), s_ptr))
        else:
            ws = self._ws(base)
            coords = torch.empty((C_, N, 3), dtype=torch.float32, device=self.device)
            energy = torch.empty(C_, dtype=torch.float64, device=self.device)
            ok = torch.empty(C_, dtype=torch.uint8, device=self.device)
            check(self.lib.mcsce_grow(self.h, C.byref(o), C.c_void_p(ws.data_ptr()), ws.numel(),
                                      C.c_void_p(coords.data_ptr()), C.c_void_p(energy.data_ptr()),
                                      C.c_void_p(ok.data_ptr()), s_ptr))
            out.update(coords=coords, energy=energy, ok=ok)
        out["_keep"] = keep
        return out

    def pick(self, out, first_valid=False):
        """Mode reduction on the device (``mcsce_pick``): index of the lowest-energy successful conformer of a ``grow``
        result (``first_valid``: of the first successful one) and its coordinates as a (n_atoms, 3) float32 NumPy array;
        ``(-1, None)`` when no conformer succeeded.  Only the picked conformer is copied to the host."""
        torch = self.torch
        n = int(out["energy"].shape[0])
        idx = torch.empty(1, dtype=torch.int32, device=self.device)
        picked = torch.empty((self.system.n_atoms, 3), dtype=torch.float32, device=self.device)
        check(self.lib.mcsce_pick(self.h, n, C.c_void_p(out["energy"].data_ptr()), C.c_void_p(out["ok"].data_ptr()),
                                  C.c_void_p(out["coords"].data_ptr()), 1 if first_valid else 0,
                                  C.c_void_p(idx.data_ptr()), C.c_void_p(picked.data_ptr()), self._stream_ptr()))
        i = int(idx.item())
        return (i, picked.cpu().numpy()) if i >= 0 else (-1, None)

    # -- single kernels (drop-in entry points and parity tests) -------------------
    def _stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def place_trials(self, residue, chis, keep4=False, oris=None):
        """K1: chis (rows, nchi) degrees -> (rows, n_sc, 3) float32 torch tensor (``keep4``: the kernel's own
        (rows, n_sc, 4) buffer).  ``chis`` may be a float64 device tensor of MAX_CHI columns.  ``oris`` (rows, <=5) rad:
        the dihedrals the reference measured before each rotation (parity replay; NaN = template constant)."""
        torch = self.torch
        st = self.system.steps[residue]
        if torch.is_tensor(chis):
            d_ch = chis.to(device=self.device, dtype=torch.float64).contiguous()
            assert d_ch.dim() == 2 and d_ch.shape[1] == MAX_CHI
            rows = d_ch
        else:
            rows = np.zeros((len(chis), MAX_CHI))
            chis = np.asarray(chis, dtype=np.float64).reshape(len(chis), -1)
            rows[:, :chis.shape[1]] = chis[:, :MAX_CHI]
            d_ch = torch.as_tensor(rows, device=self.device)
        out = torch.empty((len(rows), st.n_sc, 4), dtype=torch.float32, device=self.device)
        d_or = None
        if oris is not None:
            o5 = np.full((len(rows), MAX_CHI_ROT), np.nan)
            oris = np.asarray(oris, dtype=np.float64).reshape(len(rows), -1)[:, :MAX_CHI_ROT]
            o5[:, :oris.shape[1]] = oris
            d_or = torch.as_tensor(o5, device=self.device)
        check(self.lib.mcsce_place_trials_replay(self.h, int(residue), len(rows), C.c_void_p(d_ch.data_ptr()),
                                                 C.c_void_p(d_or.data_ptr()) if d_or is not None else None,
                                                 C.c_void_p(out.data_ptr()), self._stream_ptr()))
        return out if keep4 else out[:, :, :3]

    def score_trials(self, residue, trial_xyz, old_xyz, env_split=0):
        """K2: trial_xyz (S, T, n_sc, 3), old_xyz (S, n_old, 3) growth order -> (S, T) float64 energies
        (+inf = clash).  S environment sets, T trials each."""
        torch = self.torch
        st = self.system.steps[residue]
        trial_xyz = torch.as_tensor(trial_xyz, dtype=torch.float32, device=self.device)
        old_xyz = torch.as_tensor(old_xyz, dtype=torch.float32, device=self.device).contiguous()
        S, T = trial_xyz.shape[0], trial_xyz.shape[1]
        assert trial_xyz.shape[2] == st.n_sc and old_xyz.shape[0] == S
        n_old = old_xyz.shape[1]
        assert n_old == st.sc_start, "environment must be the %d atoms that precede this side chain" % st.sc_start
        if trial_xyz.shape[-1] == 4 and trial_xyz.is_contiguous():
            t4 = trial_xyz                                   # K1's own buffer
        else:
            t4 = torch.zeros((S, T, st.n_sc, 4), dtype=torch.float32, device=self.device)
            t4[..., :3] = trial_xyz
        env = torch.empty((S, self.system.n_atoms, 4), dtype=torch.float32, device=self.device)
        sp = self._stream_ptr()
        check(self.lib.mcsce_pack_env(self.h, S, n_old, C.c_void_p(old_xyz.data_ptr()), C.c_void_p(env.data_ptr()), sp))
        nbytes = int(self.lib.mcsce_score_workspace_bytes(self.h, S, T))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        out = torch.empty((S, T), dtype=torch.float64, device=self.device)
        check(self.lib.mcsce_score_trials(self.h, int(residue), S, T, C.c_void_p(t4.data_ptr()), C.c_void_p(env.data_ptr()),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, int(env_split), sp))
        return out

    def select(self, energies, probs, beta, uniforms):
        """K3: energies (S, T), probs (T,), uniforms (S,) -> (selected (S,) int32, e_selected (S,) f64)."""
        torch = self.torch
        e = torch.as_tensor(energies, dtype=torch.float64, device=self.device).contiguous()
        p = torch.as_tensor(probs, dtype=torch.float64, device=self.device).contiguous()
        u = torch.as_tensor(uniforms, dtype=torch.float64, device=self.device).contiguous()
        S, T = e.shape
        sel = torch.empty(S, dtype=torch.int32, device=self.device)
        es = torch.empty(S, dtype=torch.float64, device=self.device)
        check(self.lib.mcsce_select(self.h, S, T, C.c_void_p(e.data_ptr()), C.c_void_p(p.data_ptr()), float(beta),
                                    C.c_void_p(u.data_ptr()), C.c_void_p(sel.data_ptr()), C.c_void_p(es.data_ptr()),
                                    self._stream_ptr()))
        return sel, es

    # -- step-by-step growth through the single-kernel entry points ------------------
    def initial_coords(self, n_conf):
        """(n_conf, n_atoms, 3) float32 growth-order coordinates before any residue is grown: input atoms and the
        conformer-independent GLY/ALA side chains in place, everything else parked at 1e15 (mcsce_init_env)."""
        torch = self.torch
        N = self.system.n_atoms
        env = torch.empty((1, N, 4), dtype=torch.float32, device=self.device)
        check(self.lib.mcsce_init_env(self.h, 1, C.c_void_p(env.data_ptr()), self._stream_ptr()))
        slot = torch.as_tensor(np.asarray(self.system.slot_of_atom, dtype=np.int64), device=self.device)
        return env[0, slot, :3].unsqueeze(0).repeat(int(n_conf), 1, 1).contiguous()

    def grow_stepwise(self, n_conf, beta, seed=0, noise_deg=0.5, chis=None, uniforms=None, extra_term=None,
                      record=None):
        """The growth loop of ``mcsce_grow`` driven from the host, one residue at a time, through the single-kernel
        entry points (K1 ``mcsce_place_trials``, K2 ``mcsce_score_trials``, K3 ``mcsce_select``).  It exists for
        energy terms that the fused loop does not carry: ``extra_term(step, coords, trial_xyz, finite) -> (C, T)``
        float64 is added to the energy of every trial that passed the clash filter, and
        ``extra_term(None, coords, None, None) -> (C,)`` to the starting energy - how the reference treats its
        coordinate-based terms (``efuncs_coords``, libenergy.py:806-808; side_chain_builder.py:149).

        ``chis`` (C, n_trials_total, MAX_CHI) / ``uniforms`` (C, n_res) replay given trial rows and draws; otherwise
        they come from torch's device generator seeded with ``seed`` (not the Philox streams of the fused path).
        Returns dict(coords, energy, ok) as ``grow`` does, plus ``selected`` (C, n_res)."""
        torch = self.torch
        assert self._has_backbone, "set_backbone first"
        sysm = self.system
        C_, L = int(n_conf), sysm.n_res
        gen = torch.Generator(device=self.device)
        words = [int(x) for x in np.atleast_1d(seed)]
        gen.manual_seed(int(np.random.SeedSequence(words).generate_state(1, dtype=np.uint64)[0] >> 1))
        if uniforms is None:
            uniforms = torch.rand((C_, L), dtype=torch.float64, device=self.device, generator=gen)
        else:
            uniforms = torch.as_tensor(np.asarray(uniforms, dtype=np.float64).reshape(C_, L), device=self.device)
        if chis is not None:
            chis = torch.as_tensor(np.asarray(chis, dtype=np.float64).reshape(C_, self.n_trials_total, MAX_CHI),
                                   device=self.device)
        if getattr(self, "_step_cache_key", None) != (id(sysm), sysm.trial_key):
            self._step_cache = {}
            self._step_cache_key = (id(sysm), sysm.trial_key)
        coords = self.initial_coords(C_)
        energy = torch.full((C_,), self.input_energy(), dtype=torch.float64, device=self.device)
        if extra_term is not None:
            energy = energy + extra_term(None, coords, None, None)
        alive = torch.ones(C_, dtype=torch.bool, device=self.device)
        selected = torch.full((C_, L), -2, dtype=torch.int32, device=self.device)
        rows_c = torch.arange(C_, device=self.device)
        for st in sysm.steps:
            if st.kind != "grow":
                continue
            i, T, o = st.index, len(st.prob), int(self.trial_off[st.index])
            if i not in self._step_cache:                    # library rows of the residue on the device, once per backbone
                pad = lambda a: np.pad(np.asarray(a, dtype=np.float64), ((0, 0), (0, MAX_CHI - a.shape[1])))  # noqa: E731
                self._step_cache[i] = tuple(torch.as_tensor(x, device=self.device) for x in
                                            (pad(st.chi_mean), pad(st.chi_sigma), np.asarray(st.prob, dtype=np.float64)))
            mean_d, sigma_d, prob_d = self._step_cache[i]
            if chis is not None:
                rows = chis[:, o:o + T, :].contiguous()
            else:
                z = torch.randn((2, C_, T, MAX_CHI), dtype=torch.float64, device=self.device, generator=gen)
                drawn = mean_d[None] + sigma_d[None] * z[0]
                drawn = torch.where(drawn > 180.0, drawn - 360.0, drawn)        # wrap_angles (rotamer_library.py:20-25)
                drawn = torch.where(drawn < -180.0, drawn + 360.0, drawn)
                rows = drawn + noise_deg * z[1]
                rows[:, :, st.n_chi_lib:] = 0.0
            xyz4 = self.place_trials(i, rows.reshape(C_ * T, MAX_CHI), keep4=True).reshape(C_, T, st.n_sc, 4)
            xyz = xyz4[..., :3]
            e = self.score_trials(i, xyz4, coords[:, :st.sc_start])
            if extra_term is not None:
                finite = torch.isfinite(e) & alive[:, None]
                e = torch.where(finite, e + extra_term(st, coords, xyz, finite), e)
            sel, e_sel = self.select(e, prob_d, beta, uniforms[:, i])
            if record is not None:
                record.append(dict(step=i, chis=rows.cpu().numpy(), energies=e.cpu().numpy(), selected=sel.cpu().numpy()))
            ok_now = sel >= 0
            selected[:, i] = torch.where(alive, sel, torch.full_like(sel, -1))
            pick = sel.clamp(min=0).long()
            new_xyz = xyz[rows_c, pick]
            keep = (alive & ok_now)[:, None, None]
            coords[:, st.sc_start:st.sc_start + st.n_sc] = torch.where(keep, new_xyz, coords[:, st.sc_start:st.sc_start + st.n_sc])
            energy = torch.where(alive & ok_now, energy + e_sel, energy)
            alive = alive & ok_now
        nan = torch.full_like(energy, float("nan"))
        return dict(coords=coords, energy=torch.where(alive, energy, nan), ok=alive.to(torch.uint8), selected=selected)

    def input_energy(self):
        torch = self.torch
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        check(self.lib.mcsce_input_energy(self.h, C.c_void_p(out.data_ptr()), self._stream_ptr()))
        return float(out.item())

    def counters(self, reset=False, computed=False):
        """(kernel launches, algorithmic pair evaluations[, pair evaluations executed]) since the last reset."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.mcsce_counters(self.h, C.byref(a), C.byref(b), C.byref(c), 1 if reset else 0))
        return (int(a.value), int(b.value), int(c.value)) if computed else (int(a.value), int(b.value))

    def profile_begin(self):
        check(self.lib.mcsce_profile(self.h, 1, None, None))

    def profile_end(self):
        """{kernel class: (milliseconds, launches)} for place / score_generic / score_special / select / hpmf."""
        ms = (C.c_double * 5)(); n = (C.c_int64 * 5)()
        check(self.lib.mcsce_profile(self.h, 0, ms, n))
        names = ("place_trials", "score_generic", "score_special", "select", "hpmf")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(names)}
