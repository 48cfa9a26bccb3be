#!/usr/bin/env python
"""Benchmark of the side-chain growth hot path (BASELINE.json metric: packed conformers/s and
trial atom-pair energies/s vs the FP32 peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]

One "step" = growing C conformers of the workload backbone end to end on the device (chi -> xyz,
all-pairs scoring, Boltzmann selection for every residue) with on-device RNG.  Under torchrun every
rank grows its own C conformers (conformer ids offset by rank: independent units, no collective on
the data path; scaling = weak).  Rank 0 prints ONE JSON line.

--impl reference times the CPU restatement of the reference (oracle/, a port: the reference itself
is pure Python and cannot travel to the GPU box) on all host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

FLOPS_PER_PAIR = 22          # algorithmic FP32 flops of one pair evaluation
WORKLOADS = {
    # name: (synth config, conformers per GPU, description)
    "cfg2": (2, 128, "76-residue helix backbone, 128 conformers per GPU"),
    "cfg3": (3, 2048, "300-residue 6-helix bundle backbone, 2048 conformers per GPU"),
    "cfg4": (4, 512, "150-residue backbone with PTM residues, 512 conformers per GPU"),
    "cfg5": (5, 32, "2000-residue 8-chain complex, 32 conformers per GPU"),
}
BETA = 1.0 / (300 * 0.008314462)


def build_workload(name):
    from mcsce_b200 import synth
    from mcsce_b200.core.rotamer_library import DunbrakRotamerLibrary, ptmRotamerLib
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system

    cfg, n_conf, desc = WORKLOADS[name]
    s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
    lib, ptm = DunbrakRotamerLibrary(path=synth.synthetic_library_path(0)), ptmRotamerLib()     # synthetic inputs, stated in "data"
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    return s, sysm, n_conf, desc


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle's C restatement of the reference on all host cores
# ----------------------------------------------------------------------------------------------
def cpu_baseline(s, sysm, n_steps=1, warmup=0, conformers_per_core=1):
    """conformers/s of the C port of the reference's growth loop (oracle/growth_c.c), one conformer per
    OpenMP thread on every host core - the reference's own parallelism is one conformer per
    multiprocessing worker (side_chain_builder.py:328-347).  The reference itself is pure Python/numba
    and lives in /root/reference, which does not exist on the GPU box; its setup (initialize_func_calc,
    63 s at 76 residues) is not part of the hot path and is not timed, like here."""
    from oracle.growth_c import COracle

    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    co = COracle(sysm, s.coords)
    setup_s = time.perf_counter() - t0
    n = cores * conformers_per_core
    times, ok = [], []
    for k in range(warmup + n_steps):
        t0 = time.perf_counter()
        out = co.grow(n, BETA, seed=1000 + k, n_threads=cores)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt); ok.append(float(out["ok"].mean()))
    return {"conformers_per_step": n, "cores": cores, "step_seconds": times, "setup_seconds": setup_s,
            "succeeded_fraction": float(np.mean(ok))}


_RESULT_FD = None


def emit(line):
    """The one JSON line of the run, to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def pairs_per_conformer(sysm):
    total = 0
    for st in sysm.steps:
        if st.kind != "grow":
            continue
        n_env = int(sysm.cls_active[st.index].sum())
        total += len(st.prob) * (st.n_sc * (st.n_sc - 1) // 2 + st.n_sc * n_env)
    return total


# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    s, sysm, n_conf, desc = build_workload(args.workload)
    res = cpu_baseline(s, sysm, n_steps=args.steps, warmup=args.warmup, conformers_per_core=1)
    sec = float(np.mean(res["step_seconds"]))
    value = res["conformers_per_step"] / sec
    sample = "%d conformers per step (1 per host core, OpenMP) of %s; C port of the reference loop (oracle/growth_c.c)" % (
        res["conformers_per_step"], desc)
    line = {
        "impl": "reference", "metric": "conformers_per_second", "value": value, "unit": "conformers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "conformers/s", "cores": res["cores"], "kind": "port", "sample": sample,
                         "pairs_per_second": value * pairs_per_conformer(sysm), "setup_seconds": res["setup_seconds"]},
        "e2e": {"value": value, "unit": "conformers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch

    from mcsce_b200.engine import GrowthEngine

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version/debug lines must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    s, sysm, n_conf, desc = build_workload(args.workload)
    if args.conformers:
        n_conf = args.conformers
    eng = GrowthEngine(local_rank).set_system(sysm).set_backbone(s.coords)
    conf0 = rank * n_conf

    def step(k, host=False):
        return eng.grow(n_conf, BETA, seed=1234 + k, conf_id0=conf0 + k * world * n_conf, host_out=host)

    for k in range(args.warmup):
        step(k)
    eng.counters(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ok_total = 0
    outs = []
    for k in range(args.steps):
        outs.append(step(args.warmup + k)["ok"])
    ev1.record()
    barrier()
    if sampler:
        sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches, pairs = eng.counters(reset=True)
    ok_total = int(sum(int(o.sum().item()) for o in outs))
    t = torch.tensor([ms, float(pairs), float(ok_total), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, pairs, ok_total, launches = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tsum[3])
    total_conf = n_conf * world * args.steps
    value = total_conf / (ms * 1e-3)

    # ---- end-to-end through the public call with host buffers: set_backbone (H2D of the step's inputs)
    #      + grow_host (D2H of coordinates, energies, flags), every step ----
    e2e_steps = max(1, min(args.steps, 3))
    eng.set_backbone(s.coords)
    step(9_999, host=True)               # untimed: first use of the host-buffer path allocates its pinned staging
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(e2e_steps):
        eng.set_backbone(s.coords)
        o = step(10_000 + k, host=True)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_conf * world * e2e_steps / (float(e2e_t[0]) * 1e-3)
    h2d = int(sysm.n_input * 12 + eng.n_trials_total * (6 * 8 * 2 + 8) + sysm.n_res * (4 * 4 + 8 * 8 + 12))
    d2h = int(n_conf * (sysm.n_atoms * 12 + 9))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (score_generic), CUDA events around every launch, separate pass ----
    eng.counters(reset=True)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile_begin()
    p0.record()
    step(20_000)
    p1.record()
    prof = eng.profile_end()
    prof_step_ms = p0.elapsed_time(p1)
    _, prof_pairs, prof_computed = eng.counters(reset=True, computed=True)
    k2_ms, k2_n = prof["score_generic"]
    clocks = sampler.summary() if sampler else {}
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_tflops = sm_count * 128 * 2 * sm_max * 1e6 / 1e12
    achieved = FLOPS_PER_PAIR * prof_pairs / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else 0.0
    traffic = None
    try:          # dram__bytes_read+write of one ncu --set full capture of this kernel (profiles/, per launch of THAT capture)
        traffic = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())
    except Exception:
        pass
    executed = FLOPS_PER_PAIR * prof_computed / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else 0.0
    roofline = {
        "bound": "fp32", "kernel": "score_generic_kernel", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops, "traffic": traffic,
        "peak_source": "FP32 CUDA-core peak = %d SMs x 128 lanes x 2 x %.0f MHz (sm_max_mhz of MEASURED_PEAKS.json; that file "
                       "holds no FP32 figure)" % (sm_count, sm_max),
        "flops_per_pair": FLOPS_PER_PAIR, "pairs_per_launch_avg": prof_pairs / max(1, k2_n),
        "launch_ms_avg": k2_ms / max(1, k2_n), "launches": k2_n, "share_of_step": k2_ms / prof_step_ms if prof_step_ms else None,
        "kernel_ms": {k: v[0] for k, v in prof.items()}, "step_ms_in_this_pass": prof_step_ms,
        "note": "achieved counts ALGORITHMIC pairs (the reference's P, libcalc.py:56) over the K2 launches (neighbour screen + "
                "scoring pass of a residue, one event pair). The kernel executes fewer: chi-independent atoms (HA, CB) are scored "
                "once per residue and trials that the clash-only screen rejects (+inf either way) are not scored - "
                "achieved_executed / frac_executed count the pair evaluations actually done (screen pairs at half weight) and "
                "are the figure to judge the kernel by. In this pass all kernels are serialised on one stream, so kernel_ms are "
                "per-kernel times; in the timed steps K1/K2s overlap K2 on a second stream",
        "achieved_executed": executed, "frac_executed": executed / peak_tflops,
        "frac_at_observed_clock": (achieved / (sm_count * 128 * 2 * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks.get("sm_mhz") else None,
    }

    cpu = None
    if world == 1 and not args.no_cpu:
        res = cpu_baseline(s, sysm, n_steps=1, warmup=0, conformers_per_core=1)
        sec = float(np.mean(res["step_seconds"]))
        v = res["conformers_per_step"] / sec
        cpu = {"value": v, "unit": "conformers/s", "cores": res["cores"], "kind": "port",
               "sample": "%d conformers (1 per host core, OpenMP) of the same backbone, C port of the reference loop "
                         "(oracle/growth_c.c), %.1f s" % (res["conformers_per_step"], sec),
               "pairs_per_second": v * pairs_per_conformer(sysm), "setup_seconds": res["setup_seconds"]}

    line = {
        "metric": "conformers_per_second", "value": value, "unit": "conformers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 pair terms, f64 accumulation/selection", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "conformers_per_gpu": n_conf,
                   "conformers_total_per_step": n_conf * world, "residues": sysm.n_res, "atoms": sysm.n_atoms,
                   "trials_per_conformer": int(eng.n_trials_total), "pairs_per_conformer": pairs_per_conformer(sysm),
                   "terms": list(sysm.terms), "rng": "on-device Philox4x32-10",
                   "l2_note": "per-step working set (environment + trial buffers, %.0f MB) exceeds the 126 MB L2"
                              % (n_conf * (sysm.n_atoms + 4000) * 16 / 1e6)},
        "pairs_per_second": pairs / (ms * 1e-3), "succeeded_fraction": ok_total / total_conf,
        "e2e": {"value": e2e_value, "unit": "conformers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "call": "GrowthEngine.set_backbone + grow(host_out=True) -> mcsce_grow_host"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--conformers", type=int, default=0, help="conformers per GPU per step (default: the workload's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...) is sent to stderr
    # by pointing fd 1 at fd 2 for the whole run; the line itself goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
