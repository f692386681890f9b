#!/usr/bin/env python
"""bench.py -- QTB-MD trajectory-steps/s (and NEGF w-points/s) on N B200s of one node.

Workload (BASELINE.json configs[4] per-GPU share, weak scaling): synthetic 3000-dof harmonic
junction, two phonon baths of 300 dofs with 4096-step diagonal memory kernels, nmd=8192,
1024 independent noise realisations per GPU.  One "step" = one velocity-Verlet step of every
trajectory on the GPU (md.vv, sclmd/md.py:367-411).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on host cores

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sclmd_b200 import synthetic as P  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128,
                    help="timed steps (default 128 = 8 time blocks, ~0.12 s: long against one nvidia-smi clock sample, which can stall the GPU for ~2 ms)")
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_diag", choices=["c5_diag", "c5_full", "c2"])
    ap.add_argument("--ntraj-per-gpu", type=int, default=None)
    ap.add_argument("--cpu-steps", type=int, default=4, help="steps of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads (C2 MD, NEGF)")
    ap.add_argument("--no-overlap", action="store_true", help="run the K.q GEMM on the same stream as the history-tail kernels (A/B measurement)")
    ap.add_argument("--tail-block", type=int, default=1, help="1: time-blocked history tails, tensor-pipe far pass spread over the steps (default); 0: direct, one ring pass per step; 2-5: earlier far-pass kernels (A/B)")
    ap.add_argument("--no-modal", action="store_true", help="propagate in real space (K.q GEMM every step) instead of the eigenbasis of md.setDyn (A/B measurement)")
    return ap.parse_args()


WORKLOADS = {
    # name: natoms, nc per bath, ml, nmd, dt, default ntraj/GPU, kernel kind, constraints
    "c5_diag": dict(natoms=1000, nc=300, ml=4096, nmd=8192, dt=0.25 / 0.658, ntraj=1024, kind="diag", fixed=0),
    "c5_full": dict(natoms=1000, nc=300, ml=4096, nmd=8192, dt=0.25 / 0.658, ntraj=256, kind="full", fixed=0),
    "c2": dict(natoms=201, nc=150, ml=1, nmd=4096, dt=0.25 / 0.658, ntraj=1024, kind="diag", fixed=60),
}


def build_problem(w):
    """K (PSD-projected as md.setDyn does), bath dof lists, kernels."""
    nph = 3 * w["natoms"]
    K, lam, U = P.psd_project_modes(P.spring_chain_dyn(w["natoms"], seed=5))
    w["_modes"] = (lam, U)                                   # what md.setDyn keeps next to the projected matrix (md.py:266-281)
    f = w["fixed"]
    cids = [list(range(f, f + w["nc"])), list(range(nph - f - w["nc"], nph - f))]
    cons = list(range(0, f)) + list(range(nph - f, nph)) if f else []
    if w["ml"] == 1:
        damp = 100 / 0.658211814201041                       # examples/runmd.py:44-47: efric = I/damp
        kern = [np.full((1, w["nc"]), 1.0 / damp) for _ in range(2)]
    elif w["kind"] == "diag":
        kern = [P.diag_kernel(w["ml"], w["nc"], w["dt"], seed=50 + b) for b in range(2)]
    else:
        kern = [P.full_kernel(w["ml"], w["nc"], w["dt"], seed=50 + b) for b in range(2)]
    return nph, K, cids, cons, kern


def algorithmic_bytes_per_traj_step(w):
    """history tail + ring write + noise row + q,p read/write + cur."""
    nph = 3 * w["natoms"]
    b = 0
    for _ in range(2):
        b += 8 * w["nc"] * (w["ml"] - 1) + 8 * w["nc"] + 8 * w["nc"]
    return b + 32 * nph + 8 * 2


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self, first=0):
        """summary of the samples taken from row `first` on (the timed region)"""
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = []
        reasons = set()
        for r in (self.rows[first:] or self.rows):
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()


def host_cores():
    info = {"cpu_count": os.cpu_count()}
    try:
        from threadpoolctl import threadpool_info
        blas = [t for t in threadpool_info() if t.get("user_api") == "blas"]
        if blas:
            info["blas_threads"] = blas[0].get("num_threads")
            info["blas"] = "%s %s" % (blas[0].get("internal_api"), blas[0].get("version"))
    except Exception:
        pass
    return info


def reference_steps(w, nsteps, warm):
    """The reference's algorithm (literal restatement oracle.LiteralMD: 3 force evaluations per step, physical
    history shift, ml matvecs per evaluation; md.py:367-474) on the host cores, ONE trajectory of this workload."""
    from oracle import sclmd_oracle as O     # bench.py may time the oracle as the CPU baseline
    nph, K, cids, cons, kern = build_problem(w)
    rng = np.random.default_rng(7)
    baths = []
    for b in range(2):
        k = kern[b]
        if k.ndim == 2:                       # the reference only knows [ml,nc,nc] kernels
            full = np.zeros((k.shape[0], w["nc"], w["nc"]))
            idx = np.arange(w["nc"])
            full[:, idx, idx] = k
            k = full
        baths.append(O.Bath("ph", cids[b], k, 0.01 * rng.standard_normal((w["nmd"], w["nc"])), w["dt"], w["nmd"]))
    m = O.LiteralMD(K, w["dt"], w["nmd"], baths, [cons] if cons else None)
    m.q, m.p = 0.05 * rng.standard_normal(nph), 0.02 * rng.standard_normal(nph)
    for c in cons:
        m.q[c] = m.p[c] = 0.0
    for _ in range(warm):
        m.vv()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        m.vv()
    return (time.perf_counter() - t0) / nsteps


def run_reference(args, w, rank):
    if rank != 0:
        return
    sec = reference_steps(w, args.steps, args.warmup)
    cores = host_cores()
    v = 1.0 / sec
    line = {"impl": "reference", "metric": "qtb_md_trajectory_steps_per_s", "value": v, "unit": "trajectory-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, w, 1),
            "cpu_baseline": {"value": v, "unit": "trajectory-steps/s", "cores": cores.get("blas_threads") or cores["cpu_count"],
                             "kind": "port", "sample": "1 trajectory x %d steps of the workload (reference algorithm, NumPy/BLAS); host %s"
                             % (args.steps, cores)},
            "e2e": {"value": v, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, w, ntraj_total):
    return {"workload": "%s: synthetic %d-dof harmonic junction, 2 phonon baths x %d dofs, %s memory kernels ml=%d, nmd=%d, "
                        "%d trajectories/GPU (BASELINE.json configs[4] per-GPU share)" %
                        (args.workload, 3 * w["natoms"], w["nc"], w["kind"], w["ml"], w["nmd"], w["ntraj"]),
            "nph": 3 * w["natoms"], "nc": w["nc"], "ml": w["ml"], "nmd": w["nmd"], "ntraj_total": ntraj_total,
            "parallelism": "trajectory-sharded x%d, no data-path collective; one all-reduce of heat-current sums" % args.gpus,
            "l2": "per-step working set (history rings %.1f GB/GPU) >> 126 MB L2" %
                  (2 * w["ntraj"] * w["ml"] * w["nc"] * 8 / 1e9)}


def make_engine(w, device, ntraj):
    from sclmd_b200.engine import MDEngine
    nph, K, cids, cons, kern = build_problem(w)
    eng = MDEngine(nph, ntraj, w["dt"], w["nmd"], device=device)
    eng.set_dyn(K)
    eng.set_modes(*w["_modes"])
    if cons:
        eng.set_constraint(cons)
    for b in range(2):
        eng.add_bath(cids[b], kern[b])
    return eng, nph


def fill_noise(eng, w, ntraj, seed, traj0):
    """Quantum coloured noise for every trajectory, generated ON THE DEVICE straight into the noise tables
    (Philox -> per-frequency factor -> in-house FFT; sclmd_md_generate_noise).  Bath temperatures T(1 +- delta/2),
    Debye-like friction spectrum.  Also returns a pinned host block of 32 time slabs per bath for the e2e leg."""
    import ctypes as C
    from sclmd_b200 import _lib, noise as N
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    stages = {"draws_ms": 0.0, "gemm_ms": 0.0, "transform_ms": 0.0, "factor_ms": 0.0}
    t_plan = t_gen = 0.0
    for b in range(2):
        gam = np.array([np.eye(w["nc"]) * 0.05 * np.pi / 6.0])
        ta = time.perf_counter()
        plan = N.ph_plan(gam, np.array([0.0]), 300.0 * (1.05 if b == 0 else 0.95), 0.5, w["dt"], w["nmd"], device=eng.device)
        tb = time.perf_counter()
        _lib.check(_lib.lib().sclmd_md_generate_noise(eng._h, b, plan._h, C.c_uint64(seed), int(traj0)))
        t_gen += time.perf_counter() - tb
        t_plan += tb - ta
        pr = plan.profile()
        for k in stages:
            stages[k] += pr[k]
        if b == 1:
            # the same table once more through the general path (dense factors: batched x = L xi product); bit-identical output
            os.environ["SCLMD_NOISE_NO_DIAG"] = "1"
            _lib.check(_lib.lib().sclmd_md_generate_noise(eng._h, b, plan._h, C.c_uint64(seed), int(traj0)))
            del os.environ["SCLMD_NOISE_NO_DIAG"]
            general = plan.profile()
        plan.close()
    gen_s = time.perf_counter() - t0 - (general["draws_ms"] + general["gemm_ms"] + general["transform_ms"]) * 1e-3
    dev_s = (stages["draws_ms"] + stages["gemm_ms"] + stages["transform_ms"]) * 1e-3
    nsamp = 2.0 * ntraj * w["nmd"] * w["nc"]
    fill_noise.report = {
        "what": "device noise generator, both baths: Philox draws x the diagonal factors of this workload's diagonal spectrum (a dense spectrum "
                "takes the batched TMA/stream-K DMMA product x = L xi: general_spectrum_path) -> mirrored transform (big-radix in-place FFT in "
                "shared memory) straight into the trajectory-major noise tables",
        "samples": nsamp, "device_s": dev_s, "samples_per_s_device": nsamp / dev_s if dev_s > 0 else None,
        "frac_of_hbm_roofline_16B_per_sample": (nsamp * 16 / dev_s) / (peaks()[0] * 1e9) if dev_s > 0 else None,
        "stage_ms": stages, "wall_s_incl_plan_setup_on_the_host": gen_s, "wall_s_plan_setup": t_plan, "wall_s_generate_calls": t_gen,
        "general_spectrum_path": {
            "what": "one bath regenerated with SCLMD_NOISE_NO_DIAG=1 (same series bit for bit)", "stage_ms_one_bath": general,
            "device_s_both_baths": 2e-3 * (general["draws_ms"] + general["gemm_ms"] + general["transform_ms"]),
            "gemm_tflops": 2.0 * w["nc"] ** 2 * (w["nmd"] // 2 + 1) * ntraj / (general["gemm_ms"] * 1e-3) / 1e12 if general["gemm_ms"] > 0 else None}}
    nblk = min(32, w["nmd"])
    blocks = []
    for b in range(2):
        blk = pinned((nblk, ntraj, w["nc"]))
        blk[...] = 0.01 * rng.standard_normal(blk.shape)
        blocks.append(blk)
    return blocks, gen_s


def negf_also(rank, world, local, barrier, fp64_peak_tflops=None):
    """NEGF transmission sweep of BASELINE.json configs[2] (runnegf.py shape: n = 483, Gamma on 150 + 150 dofs, damp = 0.1 ps,
    10^5 frequencies over 8 GPUs) with the frequency grid sharded over the ranks (weak scaling: 12500 points per GPU); host
    buffers in and out.  A second, profiled pass (one stream, CUDA events around every launch) gives the roofline of its
    dominant kernel, the rank-64 trailing update on the FP64 tensor pipe."""
    import ctypes as C
    from sclmd_b200 import _lib
    from sclmd_b200.negf import bpt
    from sclmd_b200 import parallel as PAR
    RPC = 6.582119569e-4
    K = P.spring_chain_dyn(201, seed=14) / RPC ** 2
    b = bpt(None, 0.25, 0.1, [list(range(60, 210)), list(range(393, 543))], [list(range(0, 60)), list(range(543, 603))],
            dynmatfile=K, num=1000, device=local)
    per = 12500
    om = np.linspace(0, 0.25 / RPC, per * world + 1)
    lo, hi = PAR.shard_range(len(om), rank, world)
    b.tm_sweep(om[lo:lo + 1184])
    b.tm_sweep(om[lo:hi])                      # warm-up at full length (workspace sized, kernels loaded)
    barrier()
    t0 = time.perf_counter()
    tm = b.tm_sweep(om[lo:hi])
    t1 = time.perf_counter()
    full = PAR.gather_blocks(tm, len(om))
    barrier()
    dt = time.perf_counter() - t0
    sweep_s = t1 - t0
    L = _lib.lib()
    dev_ms = C.c_double(0.0)
    _lib.check(L.sclmd_bpt_get_profile(b._handle(), None, None, None, C.byref(dev_ms)))
    device_s = dev_ms.value * 1e-3
    if world > 1:
        dt, sweep_s, device_s = max_over_ranks(dt), max_over_ranks(sweep_s), max_over_ranks(device_s)
    flops_w = (8 / 3) * 483 ** 3 + 8 * 483 ** 2 * 150          # one complex LU + n_R solves per frequency
    out = {"metric": "negf_omega_points_per_s", "value": len(om) / dt, "unit": "omega-points/s", "n": 483, "n_omega": len(om),
           "fp64_tflops_algorithmic": flops_w * len(om) / dt / 1e12, "transmission_checksum": float(np.sum(full)),
           "sweep_s_max_over_ranks": sweep_s, "device_s_max_over_ranks": device_s, "total_s_incl_allgather": dt,
           "device_only_omega_points_per_s": len(om) / device_s if device_s > 0 else None,
           "algorithmic_flops_per_omega": flops_w,
           "config": "BASELINE configs[2] shape: n=483, Gamma on 150+150 dofs, damp=0.1 ps, %d omega per GPU (10^5 over 8 GPUs), "
                     "blocks sharded + all-gather; host buffers in and out" % per}
    if rank == 0:
        nprof = 2960
        _lib.check(L.sclmd_bpt_set_profiling(b._handle(), 1))
        b.tm_sweep(om[lo:lo + nprof])
        ms = (C.c_double * 7)()
        n = (C.c_int64 * 7)()
        gf = C.c_double(0.0)
        _lib.check(L.sclmd_bpt_get_profile(b._handle(), ms, n, C.byref(gf), C.byref(dev_ms)))
        _lib.check(L.sclmd_bpt_set_profiling(b._handle(), 0))
        names = ["k_build", "k_panel", "k_gemm (rank-16, panel columns)", "k_block_trsm", "k_gemm (rank-64 trailing update)", "back substitution (k_block_trsm_upper + k_gemm)", "k_observe"]
        tot = sum(ms)
        peak = fp64_peak_tflops or 37.1
        ach = gf.value / (ms[4] * 1e-3) / 1e12 if ms[4] > 0 else None
        out["roofline"] = {
            "kernel": "k_gemm (complex rank-64 trailing update, 4 real DMMA.8x8x4 per fragment pair)", "bound": "tensor",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if ach else None, "traffic": None,
            "peak_source": "measured live: FP64 DMMA.8x8x4 chain probe (sclmd_probe_fp64)",
            "flops_per_launch": gf.value / max(1, n[4]), "avg_launch_ms": ms[4] / max(1, n[4]), "launches_timed": int(n[4]),
            "share_of_sweep": ms[4] / tot if tot > 0 else None, "sample": "%d frequencies, one stream, event pair per launch" % nprof,
            "kernel_ms": {names[i]: ms[i] for i in range(7)}, "kernel_launches": {names[i]: int(n[i]) for i in range(7)},
            "whole_sweep_algorithmic_frac_of_peak": (flops_w * nprof / (tot * 1e-3) / 1e12) / peak if tot > 0 else None}
    return out


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    args = parse()
    w = dict(WORKLOADS[args.workload])
    if args.ntraj_per_gpu:
        w["ntraj"] = args.ntraj_per_gpu
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    # stdout carries exactly ONE line, the JSON: native libraries that print there (NCCL's version banner) are sent to stderr
    # while the run lasts; the descriptor is restored right before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sclmd_b200 import build as _b
    if rank == 0:
        _b.build()
    if world > 1:
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ntraj = w["ntraj"]
    eng, nph = make_engine(w, local, ntraj)
    blocks, noise_gen_s = fill_noise(eng, w, ntraj, seed=1000, traj0=rank * ntraj)
    rng = np.random.default_rng(2000 + rank)
    eng.set_state(0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph)), 0)
    eng.set_tail_block(args.tail_block)
    if args.no_modal:
        eng.set_modal(False)
    if args.no_overlap:
        eng.set_overlap(False)
    K, W = args.steps, max(args.warmup, 3)

    # ---------------- device-resident throughput (`value`)
    # The time-blocked history pass works in 32-step blocks (one slice per step by default; one pass per 16 steps with --tail-block 5): the timed region always starts on a block boundary (extra untimed
    # steps), so K steps contain ceil(K/32) passes -- exact for multiples of 32, pessimistic otherwise, never optimistic.
    TBLK = 32 if args.tail_block in (1, 4) else 16      # 1: tensor-pipe far pass, 32-step blocks
    W_aligned = W + (-W) % TBLK
    # clock sampling (nvidia-smi, one sample per 100 ms) starts before the warm-up and the warm-up is extended (whole time blocks)
    # until the first sample has arrived, so that the timed region is guaranteed to contain samples taken under load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.run(W_aligned)
    if rank == 0:
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 5.0:
            eng.run(TBLK)
            W_aligned += TBLK
    if world > 1:                                            # every rank runs the same number of warm-up steps
        wa = torch.tensor([W_aligned], dtype=torch.int64, device="cuda")
        dist.all_reduce(wa, op=dist.ReduceOp.MAX)
        extra = int(wa[0]) - W_aligned
        if extra > 0:
            eng.run(extra)
            W_aligned += extra
    # Per-kernel CUDA events (the roofline block) are recorded inside the timed region when the step is a launch chain anyway
    # (baths with history tails).  A step without tails replays a CUDA graph, which per-kernel events would switch off: there
    # the timed region runs unprofiled and the same K steps are repeated once with events for the roofline block.
    # The eigenbasis step runs its products on a second stream beside the history-tail kernels: event pairs would then also contain the
    # time a kernel waits for SMs that a kernel of the other stream holds (none of them co-reside).  That case is timed unprofiled as
    # well, and the roofline block comes from a repeat of the same K steps on ONE stream, where an event pair brackets exactly one kernel.
    modal_two_streams = w["ml"] > 1 and eng.modal_active() and not args.no_overlap
    prof_in_region = w["ml"] > 1 and not modal_two_streams
    if prof_in_region:
        eng.set_profiling(True)
    l0 = eng.launch_count()
    barrier()
    n_rows0 = len(sampler.rows)
    t_wall = time.perf_counter()
    ms = eng.run(K)                                          # CUDA events on the engine's stream
    sums = np.array([eng.current_sums(b).sum() for b in range(2)] + [float(ntraj)])
    ar_ms = 0.0
    if world > 1:                                            # the one collective of the path: heat-current sums
        tsum = torch.from_numpy(sums).cuda()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(tsum)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1)
        sums = tsum.cpu().numpy()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    launches = eng.launch_count() - l0
    ms_prof = ms
    if not prof_in_region:
        if modal_two_streams:
            eng.set_overlap(False)
        eng.set_profiling(True)
        ms_prof = eng.run(K)
        if modal_two_streams:
            eng.set_overlap(True)
    prof_all = eng.profile_ex()
    modal = eng.modal_active()
    eng.set_profiling(False)
    if rank == 0 and len(sampler.rows) == n_rows0:          # a region shorter than the sampling period: take the next sample under load
        t_wait = time.perf_counter()
        while len(sampler.rows) == n_rows0 and time.perf_counter() - t_wait < 1.0:
            eng.run(TBLK)
    clocks = sampler.stop(n_rows0) if rank == 0 else None
    tot = torch.tensor([ms + ar_ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, launches_total = float(mx[0]), int(sm[1])
    else:
        ms_total, launches_total = float(tot[0]), int(tot[1])
    value = ntraj * world * K / (ms_total * 1e-3)

    # ---------------- end to end through the C ABI with host buffers (`e2e`)
    obs = pinned((3, ntraj))
    _, _, t_now = eng.get_state()
    h2d = sum(blk[0:1].nbytes for blk in blocks)
    d2h = obs.nbytes
    n_warm = 2 + (-(int(t_now) + 2)) % TBLK                  # warm, and start the timed loop on a time-block boundary
    for s in range(n_warm):
        for b in range(2):
            eng.set_noise_rows(b, (t_now + 1) % w["nmd"], blocks[b][s % 32:s % 32 + 1])
        eng.run(1)
        eng.step_observables(t_now % w["nmd"], obs)
        t_now += 1
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        for b in range(2):                                   # this step's new input: noise row t+1 of every bath
            eng.set_noise_rows(b, (t_now + 1) % w["nmd"], blocks[b][s % 32:s % 32 + 1])
        eng.run_async(1)
        eng.step_observables(t_now % w["nmd"], obs)          # this step's result: etot and heat currents (synchronises)
        t_now += 1
    torch.cuda.synchronize()                                 # the read-back only waits for evaluation A: drain the last step
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = ntraj * world * K / float(te[0])

    # the direct (one ring pass per step) history kernel, timed alone: the HBM-streaming kernel of SURVEY 8d
    direct_tail = None
    if w["ml"] > 1 and w["kind"] == "diag":
        tms = eng.time_tail(0, 5)
        hp, _ = peaks()
        gbs = 8.0 * w["nc"] * (w["ml"] - 1) * ntraj / (tms * 1e-3) / 1e9
        direct_tail = {"kernel": "k_tail_diag<4>", "avg_launch_ms": tms, "achieved_GBs": gbs, "frac_of_hbm_peak": gbs / hp}
    pf_ms = eng.time_potforce(5)
    kq_alone = {"kernel": "dgemm_tma_kernel (K.q) timed alone", "avg_launch_ms": pf_ms,
                "tflops": 2.0 * (3 * w["natoms"]) ** 2 * ntraj / (pf_ms * 1e-3) / 1e12}
    import ctypes as _C
    from sclmd_b200 import _lib as _L
    probe = {}
    for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops")):
        v = _C.c_double(0)
        _L.check(_L.lib().sclmd_probe_fp64(local, kind, _C.byref(v)))
        probe[name] = v.value
    eng.close()
    also = None
    if not args.no_also:
        also = negf_also(rank, world, local, barrier, probe["dmma_tflops"])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (CUDA-event pairs around every launch in the timed region)
    hbm_peak, peak_src = peaks()
    fp64_peak = probe["dmma_tflops"]
    nph_ = 3 * w["natoms"]
    alg_ring = 8.0 * w["nc"] * (w["ml"] - 1) * ntraj          # 8*nc*(ml-1) B per trajectory-step per bath
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_tail_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    cands = []
    pa = prof_all
    if pa["potforce"]["launches"]:
        per = pa["potforce"]["ms"] / pa["potforce"]["launches"]
        fl = 2.0 * nph_ * nph_ * ntraj
        cands.append({"kernel": "dgemm_tma_kernel (K.q: TMA-fed persistent stream-K, DMMA.8x8x4)", "bound": "tensor", "achieved": fl / (per * 1e-3) / 1e12,
                      "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / (per * 1e-3) / 1e12 / fp64_peak,
                      "traffic": (traffic or {}).get("dgemm_dram_bytes_per_launch"),
                      "peak_source": "measured live: FP64 DMMA.8x8x4 chain probe (sclmd_probe_fp64); DFMA chain %.1f" % probe["dfma_tflops"],
                      "algorithmic_flops_per_launch": fl, "avg_launch_ms": per, "launches_timed": pa["potforce"]["launches"],
                      "share_of_step": pa["potforce"]["ms"] / ms_prof})
    ncs = 2 * (w["nc"] + w["nc"] % 2)
    for key, name, K_ in (("modal_scatter", "scatter product W = (fC + fA) . E, [ntraj x sum nc] . [sum nc x nph]", ncs),
                          ("modal_gather", "gather product (K q')[cids] = Q' . (E lam)^T, [ntraj x nph] . [nph x sum nc]", nph_)):
        if pa[key]["launches"]:
            per = pa[key]["ms"] / pa[key]["launches"]
            fl = 2.0 * nph_ * ncs * ntraj
            cands.append({"kernel": "dgemm_tma_kernel (eigenbasis mode, %s; TMA-fed persistent stream-K, DMMA.8x8x4)" % name, "bound": "tensor",
                          "achieved": fl / (per * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / (per * 1e-3) / 1e12 / fp64_peak,
                          "traffic": None, "peak_source": "measured live: FP64 DMMA.8x8x4 chain probe (sclmd_probe_fp64)",
                          "algorithmic_flops_per_launch": fl, "avg_launch_ms": per, "launches_timed": pa[key]["launches"],
                          "share_of_step": pa[key]["ms"] / ms_prof})
    if pa["tail_far"]["launches"]:
        per = pa["tail_far"]["ms"] / pa["tail_far"]["launches"]
        far_fl = 2.0 * TBLK * w["nc"] * w["ml"] * ntraj            # 2 nc ml per trajectory-step and bath, TBLK steps per pass
        if args.tail_block == 1:
            # 32-step blocks on the tensor pipe, the pass for the next block worked off one slice per step (+ a short mid pass at the block
            # boundary): FP64-bound (the ring is read once per 32 steps), reported against the DMMA peak on the flops of the timed region;
            # its HBM figures ride along
            nlaunch = pa["tail_far"]["launches"]
            far_fl_region = 2.0 * (2.0 * w["nc"] * w["ml"] * ntraj) * K          # two baths, 2 nc ml per trajectory-step each
            far_ms = pa["tail_far"]["ms"]
            cands.append({"kernel": "k_tail_far_mma<32> (time-blocked history pass: Hankel x history DMMA.8x8x4 product per dof over 32-step blocks, ring "
                                    "streamed by cp.async.bulk.tensor boxes; the pass of the NEXT block runs one slice of ~148 CTAs per step and bath)",
                          "bound": "tensor",
                          "achieved": far_fl_region / (far_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": far_fl_region / (far_ms * 1e-3) / 1e12 / fp64_peak,
                          "traffic": (traffic or {}).get("far_slice2_dram_bytes_per_launch"),
                          "traffic_note": "ncu DRAM bytes of one per-step launch (one slice of each bath), profiles/r02X_farslice_ncu_raw.csv",
                          "peak_source": "measured live: FP64 DMMA.8x8x4 chain probe (sclmd_probe_fp64)",
                          "algorithmic_flops_per_launch": far_fl_region / max(1, nlaunch), "algorithmic_bytes_per_launch": 2.0 * alg_ring * K / 32.0 / max(1, nlaunch),
                          "avg_launch_ms": far_ms / max(1, nlaunch), "launches_timed": nlaunch, "ms_per_step": far_ms / K,
                          "share_of_step": far_ms / ms_prof,
                          "hbm_GBs": 2.0 * alg_ring * K / 32.0 / (far_ms * 1e-3) / 1e9, "hbm_frac_of_peak": 2.0 * alg_ring * K / 32.0 / (far_ms * 1e-3) / 1e9 / hbm_peak})
        else:
            cands.append({"kernel": ("k_tail_far_wsx<1,32,4,20,36> (time-blocked history pass, 32 steps per ring pass; producer warp + 20 bulk-copy stages of 4 ring rows)" if TBLK == 32 else
                                     "k_tail_far_ws<2,8,5> (time-blocked history pass, 16 steps per ring pass; producer warp + 5 bulk-copy stages of 8 ring rows)"), "bound": "hbm",
                          "achieved": alg_ring / (per * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": alg_ring / (per * 1e-3) / 1e9 / hbm_peak, "traffic": (traffic or {}).get("far_dram_bytes_per_launch"),
                          "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_ring, "avg_launch_ms": per,
                          "launches_timed": pa["tail_far"]["launches"], "share_of_step": pa["tail_far"]["ms"] / ms_prof,
                          "fp64_tflops": far_fl / (per * 1e-3) / 1e12, "fp64_frac_of_peak": far_fl / (per * 1e-3) / 1e12 / fp64_peak})
    if pa["tail_direct"]["launches"] and w["kind"] == "full":
        per = pa["tail_direct"]["ms"] / pa["tail_direct"]["launches"]
        fl = 2.0 * (w["ml"] - 1) * w["nc"] * w["nc"] * ntraj
        cands.append({"kernel": "dgemm_tma_kernel (full memory-kernel tail: TMA-fed stream-K DMMA contraction over the rotating history ring)", "bound": "tensor",
                      "achieved": fl / (per * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / (per * 1e-3) / 1e12 / fp64_peak,
                      "traffic": None, "peak_source": "measured live: FP64 DMMA.8x8x4 chain probe (sclmd_probe_fp64)",
                      "algorithmic_flops_per_launch": fl, "avg_launch_ms": per, "launches_timed": pa["tail_direct"]["launches"],
                      "share_of_step": pa["tail_direct"]["ms"] / ms_prof})
    elif pa["tail_direct"]["launches"]:
        per = pa["tail_direct"]["ms"] / pa["tail_direct"]["launches"]
        cands.append({"kernel": "k_tail_diag<4> (direct history pass, every step)", "bound": "hbm",
                      "achieved": alg_ring / (per * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": alg_ring / (per * 1e-3) / 1e9 / hbm_peak, "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_ring, "avg_launch_ms": per,
                      "launches_timed": pa["tail_direct"]["launches"], "share_of_step": pa["tail_direct"]["ms"] / ms_prof})
    cands.sort(key=lambda c: -c["share_of_step"])
    roof = cands[0] if cands else None
    if roof is not None:
        roof["other_kernels"] = cands[1:]
        roof["elementwise_kernels_ms_per_step"] = {k: pa[k]["ms"] / K for k in ("modal_bath", "modal_update", "tail_near") if pa[k]["launches"]}
        tails_fl = sum(2.0 * w["nc"] * w["ml"] for _ in range(2)) * ntraj if w["kind"] == "diag" else sum(2.0 * w["nc"] ** 2 * w["ml"] for _ in range(2)) * ntraj
        exec_fl = (2.0 * 2.0 * nph_ * ncs * ntraj if modal else 2.0 * nph_ * nph_ * ntraj) + tails_fl
        roof["whole_step"] = {"executed_flops_per_step": exec_fl, "executed_tflops": exec_fl / (ms / K * 1e-3) / 1e12,
                              "frac_of_fp64_peak": exec_fl / (ms / K * 1e-3) / 1e12 / fp64_peak,
                              "algorithmic_flops_per_step_real_space": 2.0 * nph_ * nph_ * ntraj + tails_fl,
                              "note": "executed = what the kernels of this mode compute (eigenbasis: gather + scatter products 4 nph sum(nc) per "
                                      "trajectory instead of K.q, 2 nph^2); the history tails are the same in both modes"}
        roof["step_algorithmic_GBs_direct_algorithm"] = algorithmic_bytes_per_traj_step(w) * ntraj * K / (ms * 1e-3) / 1e9
        roof["direct_tail_kernel_standalone"] = direct_tail
        kq_alone["frac_of_fp64_peak"] = kq_alone["tflops"] / fp64_peak
        if modal:
            kq_alone["note"] = "not on the path in the eigenbasis mode; kept as the measurement of the real-space K.q product"
        else:
            kq_alone["share_of_step_serialised"] = kq_alone["avg_launch_ms"] * K / ms     # what an ncu launch list (serialised kernels) shows
        roof["kq_gemm_standalone"] = kq_alone
        roof["profiled_pass"] = ("the timed region itself" if prof_in_region else
                                 ("a repeat of the same K steps on one stream with an event pair per launch (%.4f ms per step: exclusive kernel times, "
                                  "the near passes are not hidden beside the products; the timed region runs the same kernels on two streams)" % (ms_prof / K))
                                 if modal_two_streams else
                                 "a repeat of the same K steps with per-kernel events (%.4f ms per step; the timed region replays CUDA graphs)" % (ms_prof / K))
        roof["note"] = ("per-kernel times are CUDA-event pairs around every launch of the profiled pass, on the stream of the launch (see "
                        "profiled_pass); share_of_step is relative to that pass and is the figure to compare with the serialised ncu launch "
                        "list under profiles/ (a step is the sum of its kernels, none of them co-reside on an SM); kq_gemm_standalone is "
                        "the real-space K.q product timed alone")
    line = {"metric": "qtb_md_trajectory_steps_per_s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "warmup_run": W_aligned, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, w, ntraj * world),
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                                      "d2h_bytes_per_step": d2h, "api": "sclmd_md_set_noise_rows (async H2D from pinned memory) + sclmd_md_run(1) + "
                                      "sclmd_md_get_step_observables (D2H, synchronises) every step"},
            "gpu_launches": launches_total, "roofline": roof, "wall_ms_timed_region": wall_ms, "allreduce_ms": ar_ms,
            "heat_current_mean": [float(sums[b] / sums[2] / w["nmd"]) for b in range(2)],
            "noise_generation_s": noise_gen_s, "noise": getattr(fill_noise, "report", None), "fp64_probe_tflops": probe, "also": also,
            "tail_mode": "time-blocked (ring streamed once per 32 steps by the tensor-pipe far pass, one slice per step)" if args.tail_block == 1 else ("time-blocked, other kernel (--tail-block %d)" % args.tail_block if args.tail_block else "direct (ring streamed every step)"),
            "propagation": ("eigenbasis of md.setDyn (sclmd_md_set_modes): diagonal harmonic force, gather + scatter products over the bath dofs"
                            if modal else "real space: K.q GEMM every step")}

    # ---------------- BASELINE configs[4], full-kernel variant (the FP64-bound one): 256 trajectories, ring-segment GEMM tails
    if world == 1 and not args.no_also and args.workload == "c5_diag":
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "c5_full", "--steps", "8", "--warmup", "3",
                                "--no-also", "--no-cpu-baseline"], capture_output=True, text=True, timeout=900)
            cf = json.loads(r.stdout.strip().splitlines()[-1])
            line["also_md_c5_full"] = {k: cf[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "config")}
            rf = cf.get("roofline") or {}
            line["also_md_c5_full"]["roofline"] = {k: rf.get(k) for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "share_of_step")}
        except Exception as exc:
            line["also_md_c5_full"] = {"error": str(exc)[:200]}

    # flat copies of the secondary headline numbers inside `roofline` (the driver keeps the scalar keys of that object)
    if roof is not None:
        if also:
            roof["negf_omega_points_per_s"] = also.get("value")
            roof["negf_whole_sweep_frac_of_fp64_peak"] = (also.get("roofline") or {}).get("whole_sweep_algorithmic_frac_of_peak")
            roof["negf_rank64_update_frac_of_fp64_peak"] = (also.get("roofline") or {}).get("frac")
        cfull = line.get("also_md_c5_full") or {}
        if "value" in cfull:
            roof["c5_full_trajectory_steps_per_s"] = cfull["value"]
            roof["c5_full_ring_gemm_frac_of_fp64_peak"] = (cfull.get("roofline") or {}).get("frac")
        nz = line.get("noise") or {}
        roof["noise_samples_per_s"] = nz.get("samples_per_s_device")
        roof["noise_frac_of_hbm_roofline"] = nz.get("frac_of_hbm_roofline_16B_per_sample")
        roof["whole_step_frac_of_fp64_peak"] = (roof.get("whole_step") or {}).get("frac_of_fp64_peak")
        roof["traffic_source"] = "stored ncu --set full capture (profiles/r02_tail_traffic.json <- profiles/r02X_farslice_ncu_raw.csv), not measured in this run"

    # ---------------- BASELINE configs[1] shape (603-dof junction of the example, ml = 1 baths, fixed ends, 1024 trajectories)
    if world == 1 and not args.no_also and args.workload != "c2":
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "c2", "--steps", "1024", "--warmup", "16",
                                "--no-also", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
            c2 = json.loads(r.stdout.strip().splitlines()[-1])
            line["also_md_config2"] = {k: c2[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "config")}
            line["also_md_config2"]["kq_gemm_tflops_in_step"] = (c2.get("roofline") or {}).get("achieved")
        except Exception as exc:  # the headline line must not depend on the secondary workload
            line["also_md_config2"] = {"error": str(exc)[:200]}

    # ---------------- BASELINE configs[3] shape: the current-induced junction (three electron baths, one biased with dense matrices)
    if world == 1 and not args.no_also:
        try:
            r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "probe_c4_md.py"), "1024", "512"],
                               capture_output=True, text=True, timeout=600)
            line["also_md_config4"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as exc:
            line["also_md_config4"] = {"error": str(exc)[:200]}

    # ---------------- BASELINE configs[0] shape: ONE trajectory of the same junction (the reference's own mode of operation)
    if world == 1 and not args.no_also:
        try:
            w1 = dict(WORKLOADS["c2"])
            e1, nph1 = make_engine(w1, local, 1)
            r1 = np.random.default_rng(3)
            for b in range(2):
                e1.set_noise(b, 0.01 * r1.standard_normal((1, w1["nmd"], w1["nc"])))
            e1.set_state(0.05 * r1.standard_normal((1, nph1)), 0.02 * r1.standard_normal((1, nph1)), 0)
            e1.run(64)
            ms1 = e1.run(4096)
            e1.set_persistent(False)
            e1.run(64)
            ms1c = e1.run(1024)
            e1.close()
            line["also_md_config1"] = {"metric": "qtb_md_trajectory_steps_per_s", "value": 4096 / (ms1 * 1e-3), "unit": "trajectory-steps/s",
                                       "us_per_step": ms1 / 4096 * 1e3, "steps": 4096,
                                       "kernel": "k_md_persist (one cooperative launch per run, one grid barrier per step)",
                                       "launch_chain_value": 1024 / (ms1c * 1e-3),
                                       "config": "BASELINE configs[0] shape: 603-dof junction, 2 time-local baths x 150 dofs, fixed ends, 1 trajectory"}
        except Exception as exc:
            line["also_md_config1"] = {"error": str(exc)[:200]}

    # ---------------- CPU baseline (reported, not the target): rank 0, N=1 only
    if world == 1 and not args.no_cpu_baseline:
        sec = reference_steps(w, args.cpu_steps, 1)
        cores = host_cores()
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "trajectory-steps/s",
                                "cores": cores.get("blas_threads") or cores["cpu_count"], "kind": "port",
                                "sample": "1 trajectory x %d steps of the same workload, reference algorithm "
                                          "(oracle.LiteralMD, NumPy/BLAS); host %s" % (args.cpu_steps, cores)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
