#!/usr/bin/env python3
"""bench.py -- BRDF evaluations/s of the GORT hot path on N B200s (BASELINE.json metric).

Workload (config.workload = "c2"): BASELINE.json configs[1], the hemispherical BRDF sweep --
18 view zeniths x 18 sun zeniths x 36 azimuths = 11 664 input lines x 2101 wavelengths
(400-2500 nm at 1 nm), one forest (LAI 4) per GPU = 24 506 064 evaluations per GPU per step.
One evaluation = one iteration of the reference's wavelength loop (gortt.c:460-567).  With N > 1 every
rank owns its own forest's sweep (weak scaling, no data-path collective: SURVEY.md 8e).

A step = one pass of the hot path (gort_brdf_batch: angle preparation, viewed proportions, hotspot,
per-wavelength loop) over the whole sweep.  The gap-probability LUT and the PROSPECT-D / Price spectra
are computed once on the GPU during set-up, exactly as the reference computes them once per run
(gortt.c:108-120, :224-227) and as the CPU baseline receives them.

    value     device-resident: inputs already in HBM, gort_brdf_batch_dev, CUDA events on the launching
              stream around the K steps, max over ranks.
    e2e       the same K steps through the host-pointer C ABI call gort_brdf_batch: pinned host buffers,
              H2D of all inputs and D2H of rsurf inside the timed region.
    roofline  the dominant kernel (rsurf_wide_kernel), timed live with CUDA events inside the library
              (gort_profile_begin/end) over the timed region.  bound = hbm: with the (sun, lambda) terms
              cached across the 36 azimuths the kernel writes 8 B per evaluation and needs ~15 FP64
              instructions per evaluation, so HBM write bandwidth binds.  The FP64 view
              (92 algorithmic flop per evaluation against the nominal 37.2 TFLOP/s and against a
              same-run DFMA microbenchmark) is reported next to it.
    cpu_baseline  the unmodified reference compiled into oracle/_ref (kind "reference"; the oracle
              restatement, kind "port", if _ref did not travel), in-process, one process per host
              core, on a bounded sample of the same sweep.

--impl reference times only that CPU arm, on all host cores, and prints the same JSON line shape.
Timed outputs are 196 MB per step (> the 126 MB L2), so every step streams to HBM; no explicit flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "BRDF evals/sec (geom x lambda x member)"
UNIT = "evals/s"
F_LAMBDA = 92.0          # algorithmic FP64 ops per evaluation
F_GEOM = 350.0           # per (line, set)
FP64_NOMINAL_TFLOPS = 37.2


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm_sorted = sorted(sm)
        # median of the upper half = clocks while the GPU was busy (idle samples drag a plain median down)
        busy = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": (busy[len(busy) // 2] if busy else None), "sm_max_mhz": (max(mx) if mx else None),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(torch, local_rank):
    """Pin this process to the CPUs of the GPU's NUMA node before any pinned host buffer is allocated: the e2e leg
    is PCIe-bound (196 MB of D2H per step) and a buffer on the far socket costs ~20 % of the link rate.
    Best effort; returns a description for the JSON line."""
    info = {"gpu_node": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(Path("/sys/bus/pci/devices/%s/numa_node" % bdf).read_text().strip())
        info["gpu_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in Path("/sys/devices/system/node/node%d/cpulist" % node).read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["cpus"] = len(cpus)
    except Exception as e:                      # noqa: BLE001 -- never fail the bench over placement
        info["error"] = str(e)[:80]
    return info


def sweep_inputs(rank):
    from gort_b200 import workloads as wk
    w = wk.c2_hemisphere(sets=1, lai0=4.0 + 0.25 * rank)
    return w


# ------------------------------------------------------------------------------------------------
# CPU arm (reference compiled into oracle/_ref, else the oracle port)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, st, lut, ang, rl, tl, rs, reps = args
    import checkers
    chk = checkers.ref() if kind == "reference" else checkers.oracle()
    t0 = time.perf_counter()
    n = chk.brdf_repeat(st, lut, ang, rl, tl, rs, reps)
    return n, time.perf_counter() - t0


def cpu_arm_setup():
    import checkers
    kind = "reference" if checkers.ref() is not None else "port"
    chk = checkers.ref() if kind == "reference" else checkers.oracle()
    w = sweep_inputs(0)
    st = w["structure"][:, 0].copy()
    lut = chk.lut(st)
    rl, tl, rs = chk.spectra(w["leaf"][:, 0], w["soil"][:, 0], w["wavelength"])
    return kind, st, lut, w["angles"], rl, tl, rs


def cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, lines_per_core, reps=1):
    """One bounded sample: every core evaluates `lines_per_core` lines (a strided slice of the sweep,
    all 2101 bands) `reps` times.  Returns (evaluations, wall seconds)."""
    jobs = []
    G = ang.shape[1]
    for c in range(cores):
        idx = (np.arange(lines_per_core) * 37 + c * 101) % G       # spread over the whole hemisphere
        jobs.append((kind, st, lut, np.ascontiguousarray(ang[:, idx].T), rl, tl, rs, reps))
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on all host cores."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind, st, lut, ang, rl, tl, rs = cpu_arm_setup()
    lines = 96          # x 2101 bands = 2.0e5 evaluations per core per step, ~0.3-0.8 s of CPU work
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, lines)
        n_tot, t_tot = 0, 0.0
        for _ in range(args.steps):
            n, t = cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, lines)
            n_tot += n; t_tot += t
    v = n_tot / t_tot
    sample = "%d lines x 2101 bands per core per step (strided slice of the c2 sweep), %d cores" % (lines, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c2", "lines": 11664, "wavelengths": 2101, "sets_per_gpu": 1,
                   "note": "CPU arm: bounded sample of the same sweep, LUT and spectra given"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 / C4 / C5 kernel measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gort_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GORT path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    g = gort_b200.Gort(local_rank)
    w = sweep_inputs(rank)
    st, ang, wl = w["structure"], w["angles"], w["wavelength"]
    G, W = ang.shape[1], wl.shape[0]
    evals_per_rank = G * W

    # ---- set-up on the GPU (not timed): LUT + spectra, as the reference does once per run ----
    lut = g.lut(st)
    rl, tl, rs = g.spectra(w["leaf"], w["soil"], wl)
    rl, tl, rs = rl[0].copy(), tl[0].copy(), rs[0].copy()

    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_st, d_lut, d_ang, d_rl, d_tl, d_rs = T(st), T(lut), T(ang), T(rl), T(tl), T(rs)
    pitch = ((W + 15) // 16) * 16          # device rows padded to 128 B: every warp store is sector-aligned
    d_out = torch.empty((1, G, pitch), dtype=torch.float64, device=dev)
    # the launching stream: a torch stream whose handle is passed through the C ABI, so that the
    # torch CUDA events below are recorded on the very stream the kernels run on
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream
    assert stream != 0

    step_dev = g.brdf_dev_bind(d_st, d_lut, d_ang, d_rl, d_tl, d_rs, d_out, stream=stream)
    # the K timed steps are back-to-back calls of the same shape into the same buffer: the library's documented
    # overlap mode (include/gort_b200.h) lets the geometry kernel of step i+1 run under the stores of step i
    g.set_overlap(True)

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = g.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(ts)
    h0 = time.perf_counter()
    for _ in range(args.steps):
        step_dev()
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / args.steps
    e1.record(ts)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    # per-kernel durations: the same K steps again, this time with CUDA events recorded on the launching
    # stream around each kernel (gort_profile_begin/end).  Kept out of the timed region above because an
    # event between the two launches of a step would defeat their programmatic-dependent-launch overlap.
    g.profile_begin(args.steps)
    for _ in range(args.steps):
        step_dev()
    barrier()
    geom_ms, rsurf_ms, nprof = g.profile_end()

    # ---- end-to-end through the host-pointer C ABI (pinned host buffers, copies inside) ----
    pin = lambda a: _pinned_copy(gort_b200, a)
    h_st, h_lut, h_ang, h_rl, h_tl, h_rs = pin(st), pin(lut), pin(ang), pin(rl), pin(tl), pin(rs)
    h_out = gort_b200.PinnedArray((1, G, W))
    for _ in range(args.warmup):
        g.brdf(h_st.array, h_lut.array, h_ang.array, h_rl.array, h_tl.array, h_rs.array, out=h_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.brdf(h_st.array, h_lut.array, h_ang.array, h_rl.array, h_tl.array, h_rs.array, out=h_out.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None       # sampled over both timed regions (device-resident and e2e)
    checksum = float(h_out.array[0, ::997, ::211].sum())

    # ---- max over ranks ----
    tt = torch.tensor([ms_total, e2e_s * 1e3, rsurf_ms, geom_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, rsurf_ms, geom_ms = [float(x) for x in tt.tolist()]

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * evals_per_rank / (ms_per_step * 1e-3)
        e2e_value = world * evals_per_rank / (e2e_ms * 1e-3 / args.steps)
        hbm_peak, peak_src = load_peaks()
        alg_bytes = 8.0 * evals_per_rank                      # rsurf only; inputs amortise to < 0.1 B/eval
        # In the timed region consecutive launches overlap (the geometry kernel of step i+1 runs under the stores
        # of step i), so the kernel's per-launch duration there is bounded above by the whole step: achieved is
        # algorithmic bytes / (CUDA-event time of the K steps / K).  The isolated duration (second pass, events
        # around each kernel, no overlap) is reported next to it.
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        achieved_isolated = alg_bytes / (rsurf_ms * 1e-3) / 1e9
        dfma = g.dfma_peak_tflops()
        alg_flops = F_LAMBDA * evals_per_rank
        tf = alg_flops / (ms_per_step * 1e-3) / 1e12
        h2d = 8 * (st.size + lut.size + ang.size + rl.size + tl.size + rs.size)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c2", "lines": G, "wavelengths": W, "sets_per_gpu": 1,
                       "evals_per_gpu_per_step": evals_per_rank,
                       "device_row_pitch_doubles": pitch,
                       "l2": "outputs are 196 MB per step (> 126 MB L2); no explicit flush",
                       "setup_not_timed": "gap-probability LUT + PROSPECT-D/Price spectra, computed once on the GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(8 * evals_per_rank), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "numa": numa,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "rsurf_wide_kernel", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": _ncu_traffic(),
                         "peak_source": peak_src, "kernel_ms": rsurf_ms, "geom_kernel_ms": geom_ms,
                         "kernel_ms_note": "isolated launches (events around each kernel, cross-call overlap off); "
                                           "achieved/frac use ms_per_step of the overlapped timed region",
                         "achieved_isolated": achieved_isolated, "frac_isolated": achieved_isolated / hbm_peak,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "fp64": {"algorithmic_flop_per_eval": F_LAMBDA, "achieved_tflops": tf,
                                  "frac_of_nominal_37.2": tf / FP64_NOMINAL_TFLOPS,
                                  "dfma_microbench_tflops": dfma, "frac_of_dfma_microbench": tf / dfma}},
            "checksum": checksum,
        }
        if world == 1 and not args.no_extras:
            try:
                out["extras"] = extras(g, torch, dev, ts, dfma)
            except Exception as e:                     # noqa: BLE001 -- the extras must never cost the main line
                out["extras"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)          # the CPU arm uses every host core
            out["cpu_baseline"] = cpu_baseline(h_out.array[0], wl)
            out["max_rel_err"] = out["cpu_baseline"].get("parity", {}).get("max_rel_err")
        print(json.dumps(out), flush=True)

    if world > 1:
        dist.destroy_process_group()


def _time_dev(torch, ts, fn, reps=3):
    """best-of-reps CUDA-event time (ms) of fn() enqueued on stream ts, after one warm-up"""
    fn()
    ts.synchronize()
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(ts); fn(); b.record(ts); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def extras(g, torch, dev, ts, dfma_tflops):
    """Device-resident timings of the other kernels on the BASELINE.json configs they serve (full sizes), each
    against the FP64 roofline (algorithmic flop counts of SURVEY.md App. D, measured DFMA peak)."""
    from gort_b200 import workloads as wk
    from gort_b200.api import LUT_STRIDE
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    E = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    stream = ts.cuda_stream
    res = {}

    def fp64(flops, ms):
        tf = flops / (ms * 1e-3) / 1e12
        return {"bound": "fp64", "achieved": tf, "peak": dfma_tflops, "unit": "TFLOP/s", "frac": tf / dfma_tflops,
                "peak_source": "in-library DFMA microbenchmark, same run"}

    # C2 with component signatures ("-prnspec"): rsurf + C, G, T, Z = 40 B per evaluation, a quarter of the sweep
    w = wk.c2_hemisphere()
    G4 = w["angles"].shape[1] // 4
    W = w["wavelength"].shape[0]
    Wp = (W + 15) // 16 * 16
    d_st, d_ang = T(w["structure"]), T(w["angles"][:, :G4])
    d_lut = E(1, LUT_STRIDE); g.lut_dev(d_st, d_lut, stream=stream)
    d_rl, d_tl, d_rs = E(1, W), E(1, W), E(1, W)
    g.spectra_dev(T(w["leaf"]), T(w["soil"]), T(w["wavelength"]), d_rl, d_tl, d_rs, stream=stream)
    d_r, d_sc = E(1, G4, Wp), E(1, G4, Wp, 4)
    sc_ms = _time_dev(torch, ts, lambda: g.brdf_dev(d_st, d_lut, d_ang, d_rl[0], d_tl[0], d_rs[0], d_r, scomp=d_sc, stream=stream))
    hbm, _ = load_peaks()
    res["c2_prnspec"] = {"lines": G4, "wavelengths": W, "brdf_ms": sc_ms, "evals_per_s": G4 * W / (sc_ms * 1e-3),
                         "roofline": {"bound": "hbm", "achieved": 40.0 * G4 * W / (sc_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                      "frac": 40.0 * G4 * W / (sc_ms * 1e-3) / 1e9 / hbm,
                                      "note": "geometry kernel + per-wavelength kernel of one isolated call, 40 B per evaluation"}}
    del d_r, d_sc

    # C3: spectral albedo + fAPAR, 10^4 sets x 3 sun angles x 211 bands x 512 quadrature nodes
    w = wk.c3_albedo()
    M, W, S = w["structure"].shape[1], w["wavelength"].shape[0], w["angles"].shape[1]
    d_st, d_leaf, d_soil, d_wl, d_ang = T(w["structure"]), T(w["leaf"]), T(w["soil"]), T(w["wavelength"]), T(w["angles"])
    d_lut = E(M, LUT_STRIDE); d_rl, d_tl, d_rs = E(M, W), E(M, W), E(M, W)
    lut_ms = _time_dev(torch, ts, lambda: g.lut_dev(d_st, d_lut, stream=stream))
    sp_ms = _time_dev(torch, ts, lambda: g.spectra_dev(d_leaf, d_soil, d_wl, d_rl, d_tl, d_rs, stream=stream))
    d_a, d_v, d_s = E(M, S, W), E(M, S, W), E(M, S, W)
    en_ms = _time_dev(torch, ts, lambda: g.energy_dev(d_st, d_lut, d_ang, d_rl, d_tl, d_rs, d_a, d_v, d_s, stream=stream))
    evals = M * S * 512 * W
    res["c3_albedo"] = {"sets": M, "sun_angles": S, "wavelengths": W, "quadrature_nodes": 512,
                        "energy_kernel_ms": en_ms, "evals_per_s": evals / (en_ms * 1e-3),
                        "roofline": fp64(M * S * 512 * (F_GEOM + W * (F_LAMBDA + 2.0)), en_ms),
                        "lut_kernel_ms": lut_ms, "luts_per_s": M / (lut_ms * 1e-3),
                        "lut_roofline": fp64(M * 2.5e6, lut_ms),
                        "spectra_kernel_ms": sp_ms, "spectra_roofline": fp64(M * W * 130.0, sp_ms),
                        "finite_fraction": float(torch.isfinite(d_a).double().mean())}
    del d_a, d_v, d_s

    # C4: EnKF forward operator, 10^5 members x 16 geometries x 7 bands (all parameters varying)
    w = wk.c4_enkf()
    M, G, W = w["structure"].shape[1], w["angles"].shape[2], w["wavelength"].shape[0]
    d_st, d_leaf, d_soil, d_wl, d_ang = T(w["structure"]), T(w["leaf"]), T(w["soil"]), T(w["wavelength"]), T(w["angles"])
    d_lut = E(M, LUT_STRIDE); d_rl, d_tl, d_rs = E(M, W), E(M, W), E(M, W); d_out = E(M, G, W)
    lut_ms = _time_dev(torch, ts, lambda: g.lut_dev(d_st, d_lut, stream=stream), reps=2)
    sp_ms = _time_dev(torch, ts, lambda: g.spectra_dev(d_leaf, d_soil, d_wl, d_rl, d_tl, d_rs, stream=stream))
    br_ms = _time_dev(torch, ts, lambda: g.brdf_dev(d_st, d_lut, d_ang, d_rl, d_tl, d_rs, d_out, stream=stream))
    evals = M * G * W
    res["c4_enkf"] = {"members": M, "geometries": G, "bands": W, "brdf_ms": br_ms, "evals_per_s": evals / (br_ms * 1e-3),
                      "roofline": fp64(M * G * (F_GEOM + W * F_LAMBDA), br_ms),
                      "lut_kernel_ms": lut_ms, "luts_per_s": M / (lut_ms * 1e-3), "spectra_kernel_ms": sp_ms,
                      "whole_member_update_ms": lut_ms + sp_ms + br_ms,
                      "evals_per_s_including_lut_and_spectra": evals / ((lut_ms + sp_ms + br_ms) * 1e-3),
                      "finite_fraction": float(torch.isfinite(d_out).double().mean())}

    # C4a: the same ensemble with the crown structure shared and only LAI varying (favd): one crown-geometry phase
    # and one crown-count loop per LUT sub-group
    wa = wk.c4_enkf(vary_structure=False)
    d_sta = T(wa["structure"])
    luta_ms = _time_dev(torch, ts, lambda: g.lut_dev(d_sta, d_lut, stream=stream), reps=2)
    res["c4_enkf"]["lai_only_lut_kernel_ms"] = luta_ms
    res["c4_enkf"]["lai_only_luts_per_s"] = M / (luta_ms * 1e-3)

    # C5: LUT generation over the structural grid (131 072 parameter sets), one GPU's share = all of it here
    st = wk.c5_lut_grid()["structure"]
    M = st.shape[1]
    d_st = T(st); d_lut = E(M, LUT_STRIDE)
    lut_ms = _time_dev(torch, ts, lambda: g.lut_dev(d_st, d_lut, stream=stream), reps=2)
    res["c5_lut_grid"] = {"luts": M, "lut_kernel_ms": lut_ms, "luts_per_s": M / (lut_ms * 1e-3),
                          "roofline": fp64(M * 2.5e6, lut_ms), "bytes_out": M * LUT_STRIDE * 8,
                          "nan_luts": int(torch.isnan(d_lut).any(dim=1).sum())}
    return res


def _pinned_copy(gort_b200, a):
    p = gort_b200.PinnedArray(a.shape)
    p.array[...] = a
    return p


def _ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    p = ROOT / "profiles" / "rsurf_wide_dram_bytes.json"
    if p.exists():
        try:
            return json.loads(p.read_text())["dram_bytes_per_launch"]
        except Exception:
            return None
    return None


def _parity_per_band(kind, st, lut, ang, rl, tl, rs, lines, gpu_rsurf, wavelength):
    import checkers
    chk = checkers.ref() if kind == "reference" else checkers.oracle()
    idx = (np.arange(lines) * 37) % ang.shape[1]
    r_ref, _, _ = chk.brdf(st, lut, np.ascontiguousarray(ang[:, idx].T), rl, tl, rs, want_scomp=False)
    r_gpu = np.asarray(gpu_rsurf)[idx]
    nan_ok = bool(np.array_equal(np.isnan(r_gpu), np.isnan(r_ref)))
    with np.errstate(invalid="ignore"):
        rel = np.abs(r_gpu - r_ref) / np.maximum(np.abs(r_ref), 1e-12)
    per_band = np.nanmax(rel, axis=0)                                      # worst line, per wavelength
    wl = np.asarray(wavelength)
    bins = [(int(lo), float(per_band[(wl >= lo) & (wl < lo + 100)].max())) for lo in range(400, 2500, 100)]
    return {"against": kind, "what": "rsurf from GPU LUT + GPU spectra + GPU BRDF vs the CPU arm's own chain",
            "lines": lines, "bands": int(per_band.size), "tolerance": 1e-9,
            "max_rel_err": float(per_band.max()), "worst_band_nm": float(wl[int(per_band.argmax())]),
            "max_rel_err_per_100nm_from": bins, "nan_positions_coincide": nan_ok,
            "pass": bool(nan_ok and per_band.max() <= 1e-9)}


def cpu_baseline(gpu_rsurf=None, wavelength=None):
    """Bounded sample (~10-20 s of CPU work) of the same sweep on the host cores.  The leg's own outputs double as
    the parity check the metric asks for: the lines core 0 evaluates (reference LUT, reference spectra, reference
    BRDF) against the GPU path's result for the same lines, worst relative error per output band."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind, st, lut, ang, rl, tl, rs = cpu_arm_setup()
    lines = 96
    parity = None
    if gpu_rsurf is not None:
        try:
            parity = _parity_per_band(kind, st, lut, ang, rl, tl, rs, lines, gpu_rsurf, wavelength)
        except Exception as e:                     # noqa: BLE001 -- a reporting extra must not cost the bench line
            parity = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, 8)                   # warm-up
        n1, t1 = cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, lines)
        reps = int(min(40, max(1, 12.0 / max(t1, 1e-3))))
        n, t = cpu_arm_step(pool, cores, kind, st, lut, ang, rl, tl, rs, lines, reps=reps)
    out = {"value": n / t, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": "%d lines x 2101 bands x %d reps per core (strided slice of the c2 sweep), in-process, "
                     "LUT and spectra given" % (lines, reps)}
    if parity is not None:
        out["parity"] = parity
    return out


if __name__ == "__main__":
    main()
