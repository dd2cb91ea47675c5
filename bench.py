#!/usr/bin/env python3
"""Benchmark of the vpsi + rhoofr hot path (BASELINE.json metric: band-FFTs/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--mesh 192] [--states 512] [--batch 32]

One "step" = one Car-Parrinello electronic step of the hot path over all states:
rhoofr (all pairs) -> cp_grp_redist(rho) [N>1] -> V broadcast [N>1] -> vpsi (all pairs);
vofrho is excluded.  3*nstate band-FFTs per step.  Default workload = BASELINE.json
configs[3], the configuration the north_star metric is quoted on (128 H2O: 192^3 mesh, 512 states
= 256 double-packed FFTs); it fits one B200.  N>1: states sharded over GPUs per part_1d
(CP_GROUPS), total work fixed -> "strong" scaling, as the north_star target (>=6x at 8 GPUs) is.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM when the
timed region starts); `e2e` is the same step through the host-pointer C ABI (the entry points
the Fortran shim binds) with pinned host buffers, H2D/D2H inside the timed region.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/staged_oracle.c,
kind "port": the reference cannot be built in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "band-FFTs/s in vpsi+rhoofr (FP64)"
UNIT = "band-FFTs/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _measured_traffic(kernel):
    """DRAM bytes per packed pair of `kernel` from the newest committed ncu --set full capture
    (profiles/*_traffic.json, written by tools/ncu_traffic.py); None if there is none."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            k = json.load(open(path))["kernels"].get(kernel)
        except Exception:
            k = None
        if k:
            best = (float(k["dram_bytes_per_pair"]), os.path.relpath(path, ROOT))
    return best


def _byte_model(info, nstate):
    """Algorithmic bytes (SURVEY 8d / DESIGN.md): C=16 ngw, S_x=16 n1 rays, S_y=16 n1 n2 zband."""
    n1, n2, n3 = info["nr"]
    C = 16.0 * info["ngw"]
    Sx = 16.0 * n1 * info["nrays"]
    Sy = 16.0 * n1 * n2 * info["zband"]
    N8 = 8.0 * n1 * n2 * n3
    npairs = (nstate + 1) // 2
    rho = npairs * (2 * C + 2 * Sx + 2 * Sy) + N8
    vps = npairs * (6 * C + 4 * Sx + 4 * Sy) + N8
    return dict(C=C, Sx=Sx, Sy=Sy, N8=N8, rhoofr=rho, vpsi=vps, step=rho + vps)


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def has_sample(self):
        try:
            return self.p is not None and os.path.getsize(self.f.name) > 0
        except OSError:
            return False

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def _cpu_legs(n, states_seed, sample_states, probe_states=16):
    """The two CPU restatements of the reference algorithm (both "port": the reference is Fortran + FFTW +
    MPI and cannot be built in this image): oracle/staged_oracle.c (built-in mixed-radix FFT, in the style
    of mltfft_default) and oracle/staged_pocketfft.py (the same staged passes with a library FFT - pocketfft
    - in the role FFTW plays in the reference's FFTW build).  Both run on every core of the affinity mask
    (NOT OMP_NUM_THREADS: torchrun exports 1).  Returns (geo, c0, f, v, cores, name, module) of the faster
    one, decided on a short probe."""
    from oracle import cpmd_oracle as orc
    from oracle import staged, staged_pocketfft

    cores = staged_pocketfft.set_threads(0)       # sets the OpenMP team of both and pocketfft's workers
    geo = orc.make_geometry(n)
    c0, f, v = orc.synthetic_inputs(geo, sample_states, seed=states_seed)
    rates = {}
    for name, mod in (("staged_oracle.c (built-in FFT)", staged), ("staged_pocketfft.py (library FFT)", staged_pocketfft)):
        ns = min(probe_states, sample_states)
        mod.rhoofr(geo, c0[:2], f[:2], 1.0, 1.0)           # page faults, thread pools
        t0 = time.perf_counter()
        mod.rhoofr(geo, c0[:ns], f[:ns], 1.0, 1.0)
        mod.vpsi(geo, c0[:ns], np.zeros_like(c0[:ns]), f[:ns], v, 1.0)
        rates[name] = 3.0 * ns / (time.perf_counter() - t0)
    best = max(rates, key=rates.get)
    mod = staged if best.startswith("staged_oracle") else staged_pocketfft
    return geo, c0, f, v, cores, best, mod, rates


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of fftnew's staged algorithm on the host cores (the faster
    of the two ports), every core of the box, a bounded sample of the workload per step."""
    if rank != 0:
        return
    n = args.mesh
    sample_states = min(args.ref_sample_states, args.states)
    geo, c0, f, v, cores, which, mod, rates = _cpu_legs(n, 1234 + n + 7 * args.states, sample_states)
    c2 = np.zeros_like(c0)

    def step():
        mod.rhoofr(geo, c0, f, 1.0, 1.0)
        mod.vpsi(geo, c0, c2, f, v, 1.0)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 3.0 * sample_states / dt
    sample = (f"{sample_states} of {args.states} states ({(sample_states + 1) // 2} packed pairs) per step, mesh {n}^3, "
              f"rhoofr+vpsi, {which}")
    cfg = _config(args, n_gpus=args.gpus)
    cfg["reference_sample"] = sample
    line = {
        "impl": "reference", "device": "cpu", "kind": "port", "host_cores": cores,
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "probe_band_ffts_per_s": rates},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU measurement (no GPU work): the reference's algorithm restated in C / NumPy+pocketfft, "
                "not the upstream Fortran build; `value` is normalised to band-FFTs/s from the sample",
    }
    _emit(line)


_COLLECTIVES_USED = {"mode": None}


def _config(args, n_gpus):
    return {"workload": f"BASELINE.json configs[3]: 128 H2O CP-MD, {args.states} states "
                        f"({(args.states + 1) // 2} double-packed FFTs), {args.mesh}^3 mesh, dual 4",
            "mesh": args.mesh, "states": args.states, "pairs_per_batch": args.batch,
            "parallelism": f"cp_groups{n_gpus} (states sharded, rho allreduce, V broadcast)",
            "collectives": {"peer": "peer (cpb_peer_* kernels over NVLink peer memory)", "nccl": "nccl (torch.distributed)",
                            None: "none", "none": "none"}[_COLLECTIVES_USED["mode"] if n_gpus > 1 else None],
            "l2": "inputs larger than L2 (c0 block + intermediates >> 126 MB per step)"}


def cpu_baseline(args):
    """Bounded sample of the same workload on the host cores (rank 0, N=1 only): the faster of the two CPU
    ports, all cores."""
    n = args.mesh
    ns = min(args.cpu_sample_states, args.states)
    geo, c0, f, v, cores, which, mod, rates = _cpu_legs(n, 1234 + n + 7 * args.states, ns)
    c2 = np.zeros_like(c0)
    t0 = time.perf_counter()
    mod.rhoofr(geo, c0, f, 1.0, 1.0)
    mod.vpsi(geo, c0, c2, f, v, 1.0)
    dt = time.perf_counter() - t0
    return {"value": 3.0 * ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ns} of {args.states} states, mesh {n}^3, rhoofr+vpsi once ({dt:.1f} s), {which}",
            "probe_band_ffts_per_s": rates}


def reference_gpu(args):
    """The reference's OWN GPU path on this B200 (rank 0, N=1 only; bench-only, see oracle/ref_gpu_driver.cu):
    its CUDA sources compiled by nvcc where they lie, cuFFT plans per cp_cufft_utils.mod.F90:336-372, stage
    order and host round trips of fftcu_methods.mod.F90 with the shipped use_cpu_unpack_x2y = use_cpu_pack_y2x
    = .TRUE., one task / device / stream, wavefunctions resident on the device (cp_cuwfn).  A bounded
    sample of the workload; `device_scatter` is the variant behind use_cpu_* = .FALSE."""
    from oracle import cpmd_oracle as orc
    from oracle import ref_gpu

    if ref_gpu.load() is None:
        return {"unavailable": "oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)"}
    n = args.mesh
    ns = min(args.ref_gpu_sample_states, args.states)
    geo = orc.make_geometry(n)
    c0, f, v = orc.synthetic_inputs(geo, ns, seed=1234 + n + 7 * args.states)
    ref = ref_gpu.RefGpu(geo, 1.0, 1.0)
    out = {"unit": UNIT, "kind": "reference code: src/cuuser_utils.cu + cuuser_utils_kernels.cu + cuFFT, stage order of "
                                 "fftcu_methods.mod.F90 (4 host<->device copies per 3-D FFT), 1 device, 1 stream",
           "sample": f"{ns} of {args.states} states, mesh {n}^3, rhoofr+vpsi"}
    for key, scatter in (("value", False), ("value_device_scatter", True)):
        ref.rhoofr(c0[:2], f[:2], device_scatter=scatter)             # plans, page faults
        t0 = time.perf_counter()
        rho = ref.rhoofr(c0, f, device_scatter=scatter)
        c2 = ref.vpsi(c0, np.zeros_like(c0), f, v, device_scatter=scatter)
        dt = time.perf_counter() - t0
        out[key] = 3.0 * ns / dt
        npts = float(n) ** 3
        out["charge_identity_ok"] = bool(abs(rho.sum() / npts - 2.0 * ns) < 1e-8 * ns)
        del c2
    ref.close()
    return out


def widened_rows(args, dev, plan, c0, f_block, v, timed):
    """Measurements for the SURVEY 8(f) rows built next to the hot path (N=1 only; device-resident,
    CUDA events, 3 warm-up + 5 timed calls each): the local part of vofrho on the density cutoff,
    one k-point of rhoofr_c + vpsi's k-branch, tauofr + vtaupsi.  Reported beside, never inside, `value`."""
    import torch

    from cpmd_b200 import Plan, gvec

    out = {}
    n = args.mesh
    nst = min(64, c0.shape[0])
    # ---- vofrho_local: rho -> rho(G) -> ppener -> V(r) on a plan built from the nhg list
    inyh_d, hg_d = gvec.half_sphere((n, n, n), (n / 2.0) ** 2)
    dplan = Plan((n, n, n), inyh_d, hg_d, device=dev.index or 0, max_batch=1)
    g = torch.Generator(device="cpu").manual_seed(7)
    nhg = dplan.ngw
    scg = torch.zeros(nhg, dtype=torch.float64)
    scg[1:] = 4.0 * np.pi / torch.from_numpy(hg_d[1:])
    eivps = torch.complex(torch.randn(nhg, generator=g, dtype=torch.float64), torch.randn(nhg, generator=g, dtype=torch.float64))
    eirop = 0.1 * torch.complex(torch.randn(nhg, generator=g, dtype=torch.float64), torch.randn(nhg, generator=g, dtype=torch.float64))
    scg, eivps, eirop = scg.to(dev), eivps.to(dev), eirop.to(dev)
    rho = torch.rand(dplan.nnr1, dtype=torch.float64, device=dev)
    vloc = torch.empty_like(rho)
    l0 = dplan.launch_count
    ms = timed(lambda: dplan.vofrho_local_dev(rho, scg, eivps, eirop, vloc), 5, 3)
    dense_bytes = 2 * (8.0 * n ** 3 + 2 * 16.0 * n * n * dplan.info["zband"] + 2 * 16.0 * n * dplan.info["nrays"] + 16.0 * nhg)
    out["vofrho_local"] = {"ms_per_call": ms, "nhg": nhg, "launches_per_call": (dplan.launch_count - l0) // 8,
                           "algorithmic_GB_per_s": dense_bytes / (ms * 1e-3) / 1e9,
                           "note": "dense forward + ppener + dense inverse on the density-cutoff plan (DESIGN 3b)"}
    del dplan, rho, vloc
    # ---- k-points: complex states [c(+G), c(-G)], one k-point
    ngw = plan.ngw
    ck = torch.cat([c0[:nst], torch.flip(c0[:nst], dims=[0]).conj()], dim=1).contiguous()
    ck[:, ngw] = 0
    hg = torch.from_numpy(np.ascontiguousarray(gvec.half_sphere((n, n, n))[1])).to(dev)
    hgkp, hgkm = hg + 0.3, hg + 0.1
    fk = np.full(nst, 2.0)
    rhok = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    c2k = torch.zeros_like(ck)

    def step_k():
        plan.rhoofr_kpt_dev(ck, fk, 1.0, hgkp, hgkm, rhok)
        plan.vpsi_kpt_dev(ck, c2k, fk, hgkp, hgkm, v)

    ms = timed(step_k, 5, 3)
    out["kpoints"] = {"ms_per_step": ms, "states": nst, "band_ffts_per_s": 3.0 * nst / (ms * 1e-3),
                      "note": "one k-point: rhoofr_c inner loop + vpsi k-branch, one complex state per transform (DESIGN 3c)"}
    del ck, c2k, rhok
    # ---- meta-GGA: tauofr (3 transforms per pair) + vtaupsi (6)
    nh = n // 2 + 1
    gk = torch.from_numpy(np.ascontiguousarray((gvec.half_sphere((n, n, n))[0].T - nh).astype(np.float64))).to(dev)
    tau = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    c2t = torch.zeros_like(c0[:nst])
    c0t = c0[:nst].contiguous()
    ft = np.ascontiguousarray(f_block[:nst])

    def step_tau():
        plan.tauofr_dev(c0t, ft, gk, tau)
        plan.vtaupsi_dev(c0t, c2t, ft, gk, v)

    ms = timed(step_tau, 5, 3)
    out["meta_gga"] = {"ms_per_step": ms, "states": nst, "band_ffts_per_s": 9.0 * nst / (ms * 1e-3),
                       "note": "tauofr + vtaupsi: 3 + 6 band-FFTs per state (DESIGN 3d)"}
    return out


def sweep_configs(args, dev, timed, peak):
    """The other BASELINE.json configurations (configs[0], [1], [2] and two points of the configs[4] sweep) on
    one GPU, device-resident, untimed in `value`: ms per CP step and the whole-step HBM roofline fraction
    (algorithmic bytes of SURVEY 8d / measured copy rate)."""
    import torch

    from cpmd_b200 import Plan, synthetic

    out = []
    for label, n, ns in (("configs[0] single H2O, 72^3 x 4 states", 72, 4),
                         ("configs[1] 64-atom Si, 96^3 x 128 states", 96, 128),
                         ("configs[2] 32 H2O, 120^3 x 128 states", 120, 128),
                         ("configs[4] sweep 256^3 x 64 states", 256, 64),
                         ("configs[4] sweep 320^3 x 32 states", 320, 32)):
        d = synthetic.make_inputs(n, ns)
        # pairs per batch: the library's own choice (about 3 GB of work space: 64 pairs up to 120^3, 16 at 256^3);
        # 320^3 runs its 16 pairs as one batch
        plan = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], device=dev.index or 0,
                    max_batch=16 if n >= 320 else 0)
        c0 = torch.from_numpy(d["c0"]).to(dev)
        v = torch.from_numpy(d["vpot"]).to(dev)
        rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
        c2 = torch.zeros_like(c0)
        f = d["f"]
        res = {}

        def step():
            res["s"] = plan.rhoofr_dev(c0, f, rho)
            plan.vpsi_dev(c0, c2, f, v)

        ms = timed(step, 5, 3)
        ek, rg, rr = res["s"]
        bm = _byte_model(plan.info, ns)
        out.append({"workload": label, "mesh": n, "states": ns, "pairs_per_batch": plan.info["max_batch"],
                    "ms_per_step": ms,
                    "band_ffts_per_s": 3.0 * ns / (ms * 1e-3),
                    "step_algorithmic_GB": bm["step"] / 1e9,
                    "step_frac": bm["step"] / (ms * 1e-3) / 1e9 / peak,
                    "charge_identity_ok": bool(abs(rg - rr) < 1e-10 * max(rg, 1e-300))})
        del plan, c0, c2, rho, v, d
        torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist

    from cpmd_b200 import Plan, dist as cdist, lib, synthetic

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, nstate = args.mesh, args.states
    first, cnt = cdist.state_block(nstate, rank, world)

    # ---- synthetic inputs: every rank generates the same data, keeps its block of states
    d = synthetic.make_inputs(n, nstate)
    f = d["f"]
    plan = Plan(d["nr"], d["inyh"], d["hg"], d["tpiba2"], d["omega"], device=local, max_batch=args.batch)
    info = plan.info
    c0_block_host = torch.from_numpy(d["c0"][first:first + cnt]).pin_memory()
    c2_block_host = torch.zeros_like(c0_block_host).pin_memory()
    v_host = torch.from_numpy(d["vpot"]).pin_memory()
    rho_host = torch.empty(plan.nnr1, dtype=torch.float64).pin_memory()
    del d["c0"]
    f_block = np.ascontiguousarray(f[first:first + cnt])

    c0 = c0_block_host.to(dev)
    c2 = torch.zeros_like(c0)
    # N > 1: rho and V live in an NVLink-mapped segment and the two collectives of the step are the
    # library's own peer-memory kernels (cpb_peer_*); CPB_COLLECTIVES=nccl selects torch.distributed
    collectives = os.environ.get("CPB_COLLECTIVES", "peer") if world > 1 else "none"
    seg = None
    nn = plan.nnr1 + (plan.nnr1 & 1)
    if collectives == "peer":
        try:
            seg = cdist.PeerSegment(2 * nn, rank, world, device=local)
        except RuntimeError as e:      # raised on every rank alike: CUDA IPC / peer access unavailable
            seg = None
            collectives = "nccl"
            if rank == 0:
                print(f"bench.py: peer-memory collectives unavailable ({e}); using torch.distributed", file=sys.stderr)
    if seg is not None:
        rho = seg.tensor(0, plan.nnr1)
        v = seg.tensor(nn, plan.nnr1)
        v.zero_()
        if rank == 0:
            v.copy_(v_host, non_blocking=False)
    else:
        v = v_host.to(dev) if rank == 0 else torch.zeros(plan.nnr1, dtype=torch.float64, device=dev)
        rho = torch.empty(plan.nnr1, dtype=torch.float64, device=dev)
    _COLLECTIVES_USED["mode"] = collectives
    stream = torch.cuda.current_stream()
    # the exchanges run on a side stream: only the first z pass of vpsi reads V (cpb_plan_set_vpot_event), so
    # the gather and the x / y passes of vpsi's first batch overlap them (CPB_BCAST_OVERLAP=0: everything on one
    # stream).  In a CP step V derives from the group-summed rho, so the broadcast is ordered after the all-reduce.
    side = torch.cuda.Stream(device=dev, priority=-1) if seg is not None else None
    overlap_bcast = seg is not None and os.environ.get("CPB_BCAST_OVERLAP", "1") != "0"

    def step_device():
        """rhoofr on the rank's block -> cp_grp_redist(rho) + the 3 group-partial scalars -> V broadcast ->
        vpsi on the rank's block.  Returns the group-summed (ekin, rsum_g, rsum_r)."""
        # device-pointer calls with CPB_ASYNC only enqueue: the stream never drains between rhoofr and vpsi, the
        # host picks up rhoofr's three sums (cpb_rhoofr_finish) while vpsi runs
        sums = plan.rhoofr_dev(c0, f_block, rho, stream=stream, flags=lib.CPB_ASYNC)

        def rho_sums():
            return sums if sums is not None else plan.rhoofr_finish()   # None: enqueue-only call pending

        if seg is not None:
            if overlap_bcast:
                # all three exchanges go to the side stream, in the order of a CP step (V derives from the
                # group-summed rho): cp_grp_redist(rhoe) (rhoofr_utils.mod.F90:457-461) -> V broadcast -> the
                # group sum of the 2-3 scalars.  vpsi is enqueued before the host waits for the
                # scalars: its gather and x / y inverse passes need neither rho nor V and run under the
                # exchanges, only the first z pass waits for V (cpb_plan_set_vpot_event).
                done = torch.cuda.Event()
                done.record(stream)
                side.wait_event(done)
                seg.allreduce(0, nn, stream=side)
                seg.bcast(nn, nn, src=0, stream=side)
                vready = torch.cuda.Event()
                vready.record(side)
                plan.set_vpot_event(vready)
                plan.vpsi_dev(c0, c2, f_block, v, stream=stream, flags=lib.CPB_ASYNC)
                ek, rg, rr = seg.allreduce_scalars(list(rho_sums()), stream=side)   # waits for the side stream only
                return ek, rg, rr
            seg.allreduce(0, nn, stream=stream)
            seg.bcast(nn, nn, src=0, stream=stream)
            ek, rg, rr = seg.allreduce_scalars(list(rho_sums()), stream=stream)
        elif world > 1:
            ek, rg, rr = rho_sums()
            cdist.cp_grp_redist(rho)
            cdist.bcast_potential(v, src=0)
            ek, rg, rr = cdist.redist_scalars(ek, rg, rr, device=dev)
        else:
            plan.vpsi_dev(c0, c2, f_block, v, stream=stream, flags=lib.CPB_ASYNC)
            return rho_sums()
        plan.vpsi_dev(c0, c2, f_block, v, stream=stream, flags=lib.CPB_ASYNC)
        return ek, rg, rr

    def checks():
        """Parity evidence on the timed configuration itself: the reference's charge self-check
        (rhoofr_utils.mod.F90:607-635) on the group-summed density, the energy identity that links the two
        routines (-sum dotp(c0,c2) = ekin + 1/N sum V rho) with group-summed scalars, and - N > 1 -
        64-bit checksums of rho and V on every rank (the peer all-reduce promises bit-identical ranks)."""
        c2.zero_()
        ek, rg, rr = step_device()
        torch.cuda.synchronize()
        if seg is not None:
            seg.check()
        npts = float(n) ** 3
        rr_reduced = rho.sum().item() * plan.omega / npts           # from the all-reduced array itself
        w = torch.full((plan.ngw,), 2.0, dtype=torch.float64, device=dev)
        w[0] = 1.0
        occ = torch.from_numpy((f_block != 0).astype(np.float64)).to(dev)
        dot = ((c0.real * c2.real + c0.imag * c2.imag) * w).sum(dim=1)
        dot = (dot * occ).sum().item()
        vrho = (v * rho).sum().item() / npts
        sums = [int(rho.view(torch.int64).sum().item()), int(v.view(torch.int64).sum().item())]
        if world > 1:
            t = torch.tensor([dot], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            dot = t.item()
            ck = torch.tensor(sums, dtype=torch.int64, device=dev)
            allck = [torch.empty_like(ck) for _ in range(world)]
            dist.all_gather(allck, ck)
            identical = all(torch.equal(a, allck[0]) for a in allck)
        else:
            identical = None
        e_test = ek + vrho
        out = {
            "charge_identity": {"rsum_g": rg, "rsum_r_of_reduced_rho": rr_reduced, "rsum_r_from_partials": rr,
                                "ok": bool(abs(rr_reduced - rg) < 1e-10 * rg and abs(rr - rg) < 1e-10 * rg)},
            "energy_identity": {"minus_sum_dotp_c0_c2": -dot, "ekin_plus_int_V_rho": e_test,
                                "ok": bool(abs(-dot - e_test) < 1e-9 * max(1.0, abs(e_test)))},
            "ranks_bit_identical": {"rho_and_V_checksums_equal": identical,
                                    "ok": bool(identical) if world > 1 else True},
        }
        out["all_ok"] = all(v_["ok"] for v_ in out.values())
        c2.zero_()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps

    # ---- device-resident timing (value): the product configuration, no per-kernel events inside the
    # timed region
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    # nvidia-smi needs ~0.1-0.3 s before its first sample: keep the GPUs under the same load (untimed
    # steps, all ranks alike) until the sampler is running, so that short timed regions (N = 8: 5 steps
    # = 20 ms) still get clock samples taken under load
    t_pre = time.perf_counter()
    while True:
        go = 1.0 if (sampler is None or sampler.has_sample() or time.perf_counter() - t_pre > 3.0) else 0.0
        if world > 1:
            flag = torch.tensor([go], dtype=torch.float64, device=dev)
            dist.broadcast(flag, src=0)
            go = flag.item()
        if go:
            break
        step_device()
    check_result = checks()
    if not check_result["all_ok"]:
        print(f"bench.py: PARITY CHECKS FAILED on the timed configuration: {json.dumps(check_result)}", file=sys.stderr)
    l0 = plan.launch_count
    ms_step = timed(step_device, args.steps, 0)
    launches = plan.launch_count - l0
    if seg is not None:
        seg.check()                    # mandatory: a barrier timeout inside the timed region voids the number
    clocks = sampler.stop() if sampler else None
    value = 3.0 * nstate / (ms_step * 1e-3)

    # ---- the same step with the device-side REAL SPACE WFN KEEP (CPB_PSI_KEEP / CPB_PSI_REUSE): vpsi
    # starts from the y-pass output rhoofr left in HBM, i.e. nstate of the 3*nstate band-FFTs are not
    # recomputed.  Reported separately as ms per CP step; `value` above never uses it.
    def step_device_keep():
        plan.rhoofr_dev(c0, f_block, rho, stream=stream, flags=lib.CPB_PSI_KEEP)
        if seg is not None:
            seg.allreduce(0, nn, stream=stream)
            seg.bcast(nn, nn, src=0, stream=stream)
        elif world > 1:
            cdist.cp_grp_redist(rho)
            cdist.bcast_potential(v, src=0)
        plan.vpsi_dev(c0, c2, f_block, v, stream=stream, flags=lib.CPB_PSI_REUSE)

    ms_step_keep = timed(step_device_keep, max(1, min(args.steps, 3)), 1)

    # ---- per-kernel durations for the roofline: a separate pass with the kernels serialised on one
    # stream (cpb_plan_set_streams(1)) and every launch bracketed by CUDA events on that stream
    prof_steps = max(1, min(args.steps, 2))
    n_streams = info["streams"]
    plan.set_streams(1)
    plan.set_profiling(True)
    step_device()
    plan.kernel_times(reset=True)
    ms_step_serial = timed(step_device, prof_steps, 0)
    ktimes = plan.kernel_times(reset=True)
    plan.set_profiling(False)
    plan.set_streams(n_streams)

    # ---- e2e through the host-pointer C ABI (what the Fortran shim binds)
    def step_host():
        plan.rhoofr(c0_block_host, f_block, rho_host, flags=lib.CPB_C0_KEEP)
        if world > 1:
            rho.copy_(rho_host, non_blocking=True)
            if seg is not None:
                seg.allreduce(0, nn, stream=stream)
            else:
                cdist.cp_grp_redist(rho)
            rho_host.copy_(rho, non_blocking=True)
            torch.cuda.synchronize()
        plan.vpsi(c0_block_host, c2_block_host, f_block, v_host, flags=lib.CPB_C0_REUSE)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ms_e2e = timed(step_host, e2e_steps, 1)

    # the same step as the MD driver issues it: forces_driver.mod.F90:175 zeroes c2 immediately
    # before CALL vpsi (:224), so a shim at that call site may pass CPB_VPSI_OVERWRITE and skip the
    # upload of the zeros (reported separately; the headline e2e keeps the subroutine's += semantics)
    def step_host_ow():
        plan.rhoofr(c0_block_host, f_block, rho_host, flags=lib.CPB_C0_KEEP)
        if world > 1:
            rho.copy_(rho_host, non_blocking=True)
            if seg is not None:
                seg.allreduce(0, nn, stream=stream)
            else:
                cdist.cp_grp_redist(rho)
            rho_host.copy_(rho, non_blocking=True)
            torch.cuda.synchronize()
        plan.vpsi(c0_block_host, c2_block_host, f_block, v_host, flags=lib.CPB_C0_REUSE | lib.CPB_VPSI_OVERWRITE)

    ms_e2e_ow = timed(step_host_ow, e2e_steps, 1)
    blk_bytes = cnt * plan.ngw * 16
    h2d = blk_bytes * 2 + plan.nnr1 * 8          # c0 block (once, kept for vpsi) + c2 block (+=) + V
    d2h = blk_bytes + plan.nnr1 * 8              # c2 block + rho
    if world > 1:
        h2d += plan.nnr1 * 8
        d2h += plan.nnr1 * 8
    e2e_value = 3.0 * nstate / (ms_e2e * 1e-3)

    extras = None
    sweep = None
    if world == 1 and not args.no_extras:
        extras = widened_rows(args, dev, plan, c0, f_block, v, timed)
        sweep = sweep_configs(args, dev, timed, _peaks()[0])

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (largest share of the step)
    peak, peak_src = _peaks()
    bm = _byte_model(info, cnt)
    dom = max(("x_inv", "y_inv", "z_rho", "z_vpsi", "y_fwd", "x_fwd"), key=lambda k: ktimes[k][0])
    tot_ms = sum(v_[0] for v_ in ktimes.values())
    dom_ms, dom_n = ktimes[dom]
    npairs_local = (cnt + 1) // 2
    pairs_per_launch = npairs_local * prof_steps / max(dom_n, 1) * (2 if dom in ("x_inv", "y_inv") else 1)
    # algorithmic bytes per packed pair of each kernel: x_inv reads the two c0 columns
    # and writes T1; x_fwd reads T1 (its band-ray output stays in L2 for k_unpack, which is charged
    # the c0 read, the c2 read-modify-write: 4C)
    per_pair = {"x_inv": 2 * bm["C"] + bm["Sx"], "y_inv": bm["Sx"] + bm["Sy"], "z_rho": bm["Sy"],
                "z_vpsi": 2 * bm["Sy"], "y_fwd": bm["Sy"] + bm["Sx"], "x_fwd": bm["Sx"]}[dom]
    per_launch_extra = bm["N8"] if dom in ("z_rho", "z_vpsi") else 0.0
    bytes_per_launch = pairs_per_launch * per_pair + per_launch_extra
    achieved = bytes_per_launch / (dom_ms / max(dom_n, 1) * 1e-3) / 1e9
    mt = _measured_traffic("k_" + dom)
    traffic = mt[0] * pairs_per_launch if mt else None
    step_bytes = _byte_model(info, nstate)["step"] / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(args, world),
        "clocks": clocks, "checks": check_result,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e, "steps": e2e_steps,
                "api": "cpb_rhoofr(CPB_C0_KEEP) + cpb_vpsi(CPB_C0_REUSE), pinned host buffers, c2 += semantics"},
        "e2e_overwrite": {"value": 3.0 * nstate / (ms_e2e_ow * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_ow,
                          "h2d_bytes_per_step": int(h2d - blk_bytes), "d2h_bytes_per_step": int(d2h),
                          "api": "as e2e, but cpb_vpsi(CPB_VPSI_OVERWRITE): the MD call site zeroes c2 first "
                                 "(forces_driver.mod.F90:175,224)"},
        "psi_keep": {"ms_per_step": ms_step_keep,
                     "note": "device-resident step with CPB_PSI_KEEP/CPB_PSI_REUSE (the analogue of CPMD's REAL SPACE WFN "
                             "KEEP, rhoofr_utils.mod.F90:350-363): vpsi reuses rhoofr's y-pass output, "
                             f"{cnt * 56.03e6 / 2 / 1e9 if n == 192 else 0:.1f} GB cache per rank; not used by `value`"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": (mt[1] + " (dram bytes per pair x pairs per launch)") if mt else None,
                     "peak_source": peak_src,
                     "bytes_per_launch": bytes_per_launch, "avg_launch_ms": dom_ms / max(dom_n, 1),
                     "kernel_share_of_step": dom_ms / max(tot_ms, 1e-9),
                     "step_algorithmic_GB": step_bytes / 1e9,
                     "step_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                     "timing": "kernels serialised on one stream, CUDA events around every launch, "
                               f"{prof_steps} step(s); the timed `value` run uses {n_streams} batch stream(s), no events",
                     "ms_per_step_serialised": ms_step_serial,
                     "kernel_ms_per_step": {k: v_[0] / prof_steps for k, v_ in ktimes.items()}},
    }
    if extras is not None:
        line["widened_rows"] = extras
    if sweep is not None:
        line["sweep"] = sweep
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
        try:
            line["reference_gpu"] = reference_gpu(args)
        except Exception as e:      # a baseline leg must not take the bench line down
            line["reference_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries print there too (NCCL's version banner
    ignores NCCL_DEBUG_FILE): point fd 1 at stderr for the whole run and keep the real stdout aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, default=192)
    ap.add_argument("--states", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sample-states", type=int, default=96,
                    help="states per step of the --impl reference arm (bounded sample of the workload)")
    ap.add_argument("--cpu-sample-states", type=int, default=256,
                    help="states of the cpu_baseline leg (about 10-30 s of CPU work)")
    ap.add_argument("--ref-gpu-sample-states", type=int, default=32,
                    help="states of the reference_gpu leg (the reference's own cuFFT path, bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the widened-row measurements (vofrho, k-points, tau)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        from cpmd_b200 import dist as cdist
        cdist.init_from_env()
    run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
