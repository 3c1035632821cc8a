#!/usr/bin/env python
"""bench.py -- nonbonded pair-interactions/s and ms/step of the B200 nbnxm path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--scaling strong|weak] [--impl reference]

Default workload for EVERY N (1, 2, 4, 8): water_1M, the 1.03 M-atom water box of BASELINE.json configs[3], strong scaling --
the configuration the north_star's 8-GPU target is defined on; at N = 1 the line also carries a `secondary` block for
water_24k (configs[1]).  --workload ref_water_24k / ref_water_96k / ref_water_1M use the reference's own benchmark water
(liquid structure, nbnxm/benchmark/bench_coords.h) instead of the jittered lattice; --scaling weak puts N copies of the
workload side by side along x.

A step = one pass of the hot path over one set of coordinates: coordinates -> grid-ordered device layout fused
with the output clear, cluster-pair force kernel (LJ + Ewald real space, force only), force un-sort: 3 launches.
  value  : useful pair interactions (non-excluded, r < rc; counted exactly on the device) per second with
           coordinates already resident in HBM, L2 flushed between steps, CUDA-event timed on the stream the
           kernels run on, max over ranks;
  e2e    : the same metric through the public nblib-style call ForceCalculator.compute(x_host) -> f_host with
           pinned HOST buffers (H2D of x and D2H of f inside the timed region);
  roofline: the force kernel alone against the FP32 FMA roofline (the path is FP32-bound), with
           the HBM view beside it;
  search : what the metric above does not contain -- steady-state re-gridding + pair search + packing, the rolling prune,
           and the step amortised over a pair-list lifetime (nstlist = 100) with dynamic pruning;
  sustained: >= 2 s of back-to-back steps with NVML sampled every 2 ms: the SM clock an MD run of this kernel holds;
  cpu_baseline: the reference's own CPU SIMD nbnxm path (oracle/_ref, compiled from the reference sources)
           on this host's cores, bounded sample.
N > 1: atoms are split into N slabs along x (spatial domain decomposition), one rank per GPU; halo coordinates and
forces go straight into the neighbours' peer-memory windows over NVLink from inside the step's kernels
(b200nb_dd_step, gmxapi_b200/domdec.py; torch.distributed / NCCL only carries set-up data); --dd-grid 2x2x2 decomposes in
three dimensions instead.  Outside the timed region the decomposed forces are checked against a single-domain run.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nonbonded pair-interactions/s"
FLOPS_PER_PAIR = {"ewald": 66, "rf": 38}  # src/gromacs/gmxlib/nrnb.cpp:101,105 (force only)
RC = 0.9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe) through NVML, every
    2 ms from a background thread: the timed loops of the small workloads last only milliseconds, far below the start-up
    time of an `nvidia-smi -lms` child process."""

    def __init__(self, index=0, power=False):
        self.index, self.sm, self.mx, self.reasons, self.err = index, [], [], set(), None
        self.power, self.watts = power, []
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nv = pynvml
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self._nv
        names = {getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
                if self.power:
                    self.watts.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: %s" % self.err], "samples": 0}
        out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
               "samples": len(self.sm)}
        if self.power and self.watts:
            out.update(sm_mhz_min=float(np.min(self.sm)), power_w_median=float(np.median(self.watts)), power_w_max=float(np.max(self.watts)))
        return out


def sustained_block(h, s, npairs, seconds=2.0, flops_per_pair=66, device=0):
    """the burst figure (`value`: 20-200 flushed steps, a few ms of GPU time) says nothing about the SM clock
    a seconds-long MD run of this FP32-bound kernel holds.  Here: >= `seconds` of back-to-back device-resident steps (CUDA graph
    replays, no L2 flush, no host synchronisation inside a batch), NVML sampled every 2 ms on a side thread: median SM clock,
    throttle reasons, power; pairs/s over the whole interval (CUDA events) and the FP32 peak AT THE MEASURED MEDIAN CLOCK."""
    import torch
    dev = torch.device("cuda", device)
    x_dev = torch.from_numpy(s.x).to(dev).contiguous()
    f_dev = torch.zeros_like(x_dev)
    stream = torch.cuda.Stream(device=dev)
    h.set_stream(stream.cuda_stream)
    xp, fp = x_dev.data_ptr(), f_dev.data_ptr()
    for _ in range(20):
        h.step(xp, fp, 0)
    h.synchronize()
    # batch size: ~20 ms of queued work per host synchronisation
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50):
        h.step(xp, fp, 0)
    e1.record(stream)
    h.synchronize()
    t1 = e0.elapsed_time(e1) / 50.0  # ms per step
    batch = max(10, int(20.0 / max(t1, 1e-3)))
    sampler = ClockSampler(device, power=True)
    sampler.start()
    n = 0
    t_wall = time.perf_counter()
    e0.record(stream)
    prev = None
    while time.perf_counter() - t_wall < seconds:
        for _ in range(batch):
            h.step(xp, fp, 0)
        n += batch
        ev = torch.cuda.Event()
        ev.record(stream)
        if prev is not None:
            prev.synchronize()  # at most two batches queued: the GPU never runs dry, the host never runs away
        prev = ev
    e1.record(stream)
    h.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    mhz = clocks.get("sm_mhz") or 0.0
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12 if mhz else None
    pps = npairs * n / (ms * 1e-3)
    return {"seconds": ms * 1e-3, "steps": n, "ms_per_step": ms / n, "pairs_per_s": pps, "clocks": clocks,
            "fp32_peak_at_median_clock_tflops": peak,
            "step_algorithmic_frac_at_median_clock": (pps * flops_per_pair / 1e12 / peak) if peak else None,
            "note": "back-to-back graph replays of the device-resident step, L2 warm, two batches of %d steps in flight" % batch}


def ncu_capture(workload, eel):
    """What the committed `ncu --set full` capture of this workload's force kernel says (profiles/r2/traffic.json, written by
    profiles/tools/ncu_traffic.py from the captures of the round): DRAM bytes of one launch and the measured FP32-pipe / issue
    utilisation; {} when there is no capture for the workload."""
    for rnd in ("r2", "r1"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", rnd, "traffic.json")))
            if eel == "ewald" and workload in d:
                return dict(d[workload])
        except Exception:  # noqa: BLE001
            pass
    return {}


def workload_system(name, copies_along_x=1):
    """The named box, or `copies_along_x` of it side by side (weak scaling)."""
    import gmxapi_b200 as g
    gen, (nx, ny, nz) = g.systems.tiles_of(name)
    return gen(nx * copies_along_x, ny, nz)


def common_config(workload, s, npairs, eel, scaling, n_gpus):
    """The part of `config` both arms print identically (the driver compares the two dicts)."""
    return {"workload": workload if (scaling == "strong" or n_gpus <= 1) else "%d x %s along x" % (n_gpus, workload),
            "atoms": int(s.n), "useful_pairs_per_step": int(npairs), "rc": RC, "rlist": RC,
            "interaction": "LJ + " + ("Ewald real space (analytical)" if eel == "ewald" else "reaction field"),
            "flavor": "force only", "scaling": scaling if n_gpus > 1 else "single GPU",
            "l2_between_steps": "GPU arm: flushed (256 MiB write) between timed steps; CPU arm: not applicable"}


def ref_instance(s, eel, nthreads):
    from oracle import gmxref
    import gmxapi_b200 as g
    if eel == "ewald":
        kw = dict(eeltype=gmxref.EEL_EWALD_ANA, ewaldcoeff=float(np.float32(g.systems.ewald_beta(RC))))
    else:
        k, c = g.systems.rf_constants(RC)
        kw = dict(eeltype=gmxref.EEL_RF, k_rf=k, c_rf=c)
    return gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, kernel=None,
                           nthreads=nthreads, **kw)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown CPU"


def host_cores():
    """PHYSICAL cores this process may run on (BASELINE.md section 2 quotes physical cores): logical CPUs of the affinity mask
    grouped by their hyper-thread siblings; the reference arm runs one OpenMP thread per physical core."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        cpus = list(range(os.cpu_count() or 1))
    cores = set()
    for c in cpus:
        try:
            sib = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip()
        except Exception:  # noqa: BLE001
            sib = str(c)
        cores.add(sib)
    return max(1, len(cores))


def time_reference(s, eel, nthreads, steps, warmup, budget_s=20.0):
    """Times the reference CPU step (x convert + SIMD kernel + force reduction) per iteration."""
    r = ref_instance(s, eel, nthreads)
    t1 = r.time_step(False, max(warmup, 1), 1)
    n = int(max(1, min(steps, budget_s / max(t1, 1e-6))))
    t = r.time_step(False, 1, n)
    tk = r.time_kernel(False, 1, max(1, min(n, 50)))
    try:
        time_reference.last_search_ms = 1e3 * sum(r.regrid_research())  # nbnxn_put_on_grid + constructPairlist, same threads
    except Exception:  # noqa: BLE001
        time_reference.last_search_ms = None
    r.close()
    return t, tk, n


def run_reference(args):
    """--impl reference: the reference's own CPU SIMD nbnxm path on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gmxref
    mult = max(args.gpus, 1) if args.scaling == "weak" else 1
    s = workload_system(args.workload, mult)  # the same box the GPU arm decomposes over args.gpus ranks
    cores = host_cores()
    if not gmxref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgmxref_nbnxm.so missing"}))
        return
    r = ref_instance(s, args.eel, cores)
    npairs = r.pair_count()
    r.close()
    t, tk, n = time_reference(s, args.eel, cores, args.steps, args.warmup, budget_s=60.0)
    val = npairs / t
    out = {"metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
           "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference",
           "config": common_config(args.workload, s, npairs, args.eel, args.scaling, args.gpus),
           "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "reference",
                            "sample": "%d full steps (x convert + 2xMM SIMD kernel + f reduce) of %s (%d atoms), one OpenMP thread on each of "
                                      "the %d physical cores of %s; kernel-only %.3f ms" % (n, args.workload if mult <= 1 else "%d x %s" % (args.gpus, args.workload),
                                                                                             int(s.n), cores, cpu_model(), tk * 1e3)},
           "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def time_search(h, s, x_dev, reps=3):
    """Steady-state pair-search step: re-gridding + search + (fresh prune) + packing of a context that has searched before, with
    the coordinates already on the device (b200nb_put_on_grid + b200nb_build_pairlist).  CUDA events on the context's stream
    around both calls (the host stalls of the calls sit between the events, so they are inside) and the host wall clock."""
    import torch
    stream = torch.cuda.ExternalStream(int(h.stream))
    ev_ms, wall_ms = [], []
    lo, hi = np.zeros(3, np.float32), np.asarray(s.box, np.float32)
    for _ in range(reps + 1):
        h.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        h.put_on_grid(x_dev.data_ptr(), lo, hi, on_device=True)
        h.build_pairlist()
        e1.record(stream)
        h.synchronize()
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        ev_ms.append(e0.elapsed_time(e1))
    return float(np.mean(ev_ms[1:])), float(np.mean(wall_ms[1:]))  # the first repetition may still grow buffers


def search_block(args, s, coul, local_rank, step_ms, k_ms, h_static, x_dev):
    """What `value` leaves out: the pair-search step and the rolling prune, and the step amortised over a
    pair-list lifetime.  Scenario (reference defaults for a GPU run, pairlist_tuning.cpp:398-572, pairlistsets.h:100-114):
    nstlist = 100, outer list at rc + 0.15 nm, dynamically pruned to rc + 0.05 nm, rolling prune of one of nstlistPrune / 2 = 5
    parts every second step."""
    import torch
    import gmxapi_b200 as g
    nstlist, nprune = 100, 10
    rlo, rli = RC + 0.15, RC + 0.05
    static_ev, static_wall = time_search(h_static, s, x_dev)
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=coul, computeVirialAndEnergy=False, device=local_rank, epsilonRf=0.0,
                            rlistOuter=rlo, rlistInner=rli)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
    h = fc.nb
    f_dev = torch.zeros_like(x_dev)
    dyn_ev, dyn_wall = time_search(h, s, x_dev)
    stream = torch.cuda.ExternalStream(int(h.stream))
    nparts = nprune // 2
    for part in range(nparts):
        h.launch_prune(0, part, nparts)
    h.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    e0.record(stream)
    for _ in range(reps):
        for part in range(nparts):
            h.launch_prune(0, part, nparts)  # k_prune + k_pack of that part
    e1.record(stream)
    h.synchronize()
    prune_part_ms = e0.elapsed_time(e1) / (reps * nparts)
    dyn_step_ms, dyn_k_ms = h.time_step(x_dev.data_ptr(), f_dev.data_ptr(), 0, 3, max(10, min(args.steps, 50)), flush_l2=not args.no_flush)
    st = h.stats()
    fc.nb.close()
    amort = dyn_step_ms + 0.5 * prune_part_ms + dyn_ev / nstlist
    return {"search_ms": static_ev, "search_wall_ms": static_wall, "search_over_force_kernel": static_ev / k_ms,
            "search_note": "steady-state b200nb_put_on_grid + b200nb_build_pairlist, coordinates on the device, rlist = rc (the list `value` runs on)",
            "scenario": {"nstlist": nstlist, "nstlist_prune": nprune, "rolling_parts": nparts, "rlist_outer": rlo, "rlist_inner": rli},
            "search_ms_dynamic": dyn_ev, "search_wall_ms_dynamic": dyn_wall,
            "prune_ms": prune_part_ms, "prune_note": "one rolling part (k_prune + re-pack of 1/%d of the entries), every second step" % nparts,
            "ms_per_step_pruned_list": dyn_step_ms, "force_kernel_ms_pruned_list": dyn_k_ms,
            "computed_pairs_per_step_pruned_list": int(st["ntiles_packed"] * 64),
            "ms_per_step_amortised": amort,
            "amortised_overhead_frac": (amort - dyn_step_ms) / dyn_step_ms,
            "amortised_vs_static_step": amort / step_ms}


def measure_single(args, workload, local_rank, full):
    """One GPU, one workload: device-resident step (flushed), force kernel alone, end to end through the public API; with
    `full` also the search / prune / amortised block, the sustained block and the CPU baseline."""
    import torch
    import gmxapi_b200 as g
    peaks = measured_peaks()
    s = workload_system(workload)
    coul = g.CoulombType.Pme if args.eel == "ewald" else g.CoulombType.ReactionField
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=coul, computeVirialAndEnergy=False, device=local_rank, epsilonRf=0.0,
                            maxTilesPerEntry=args.max_tiles)
    t0 = time.perf_counter()
    fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
    fc.nb.synchronize()
    t_setup = time.perf_counter() - t0
    h = fc.nb
    npairs = h.pair_count(RC)
    st = h.stats()
    stream = torch.cuda.Stream(device=torch.device("cuda", local_rank))
    h.set_stream(stream.cuda_stream)  # torch-owned stream: events and the L2 flush order against the kernels

    # ---- device-resident step ------------------------------------------------------------------------------
    x_dev = torch.from_numpy(s.x).to("cuda", non_blocking=False).contiguous()
    f_dev = torch.zeros_like(x_dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if not args.no_flush else None
    torch.cuda.synchronize()
    xp, fp = x_dev.data_ptr(), f_dev.data_ptr()
    steps = args.steps

    def step():
        h.step(xp, fp, 0)  # 3 launches: x -> grid layout + output clear, force kernel, f -> atom order

    for _ in range(args.warmup):
        step()
    h.synchronize()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = h.stats()["nlaunches"]
    for k in range(steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(k & 0xff)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    h.synchronize()
    torch.cuda.synchronize()
    launches = h.stats()["nlaunches"] - l0
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))

    # ---- force kernel alone (roofline): CUDA events around its launch inside the same step, on the kernels' stream ----
    _, k_ms = h.time_step(xp, fp, 0, 3, max(10, min(steps, 100)), flush_l2=not args.no_flush)
    k_ms_cold = h.time_force_kernel(-1, 0, 3, max(10, min(steps, 50)), flush_l2=not args.no_flush)
    clocks = sampler.stop()

    # ---- end to end through the public API with pinned host buffers -----------------------------------------
    x_pin = torch.from_numpy(s.x.copy()).pin_memory()
    f_pin = torch.empty_like(x_pin).pin_memory()
    xh, fh = x_pin.numpy(), f_pin.numpy()
    for _ in range(args.warmup):
        fc.compute(xh, fh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fc.compute(xh, fh)  # synchronous: returns after the D2H of the forces completed
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    f_host = fh.copy()

    # ---- parity spot-check of what was timed (cheap, not in any timed region) -------------------------------
    assert np.allclose(f_dev.cpu().numpy(), f_host, rtol=1e-3, atol=1e-2 * np.abs(f_host).mean())

    flops = FLOPS_PER_PAIR[args.eel]
    fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    achieved = npairs * flops / (k_ms * 1e-3) / 1e12
    npad, ntiles = st["natoms_padded"], st["ntiles_packed"]
    # xq 16 + lj 8 read once, f 16 read-modify-written (32) per slot; 64 B of j-slot indices per packed tile (a warp step = 2 tiles
    # = 16 slots for each of its two half-entries); two 16 B half-entry headers per entry of the cluster-pair list
    alg_bytes = npad * (16 + 8 + 32) + ntiles * 64 + st["nentries"] * 32
    cap = ncu_capture(workload, args.eel)
    out = {
        "value": npairs / (step_ms * 1e-3), "ms_per_step": step_ms,
        "config": common_config(workload, s, npairs, args.eel, "strong", 1),
        "details": {"computed_pairs_per_step": int(ntiles * 64), "useful_lane_fraction": npairs / float(ntiles * 64),
                    "cluster_pair_lanes_before_packing": int(st["ntiles_inner"] * 64), "list_entries": int(st["nentries"]),
                    "setup_s": t_setup, "parallelism": "1 GPU", "setup": fc.nb.describe(),
                    "l2": "inputs < L2; L2 flushed (256 MiB write) between timed steps" if not args.no_flush else "not flushed"},
        "roofline": {"bound": "fp32", "kernel": "k_force<Ewald,geometric LJ,F>" if args.eel == "ewald" else "k_force<RF,geometric LJ,F>",
                     "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                     "kernel_ms": k_ms, "kernel_ms_alone_after_l2_flush": k_ms_cold, "flops_per_useful_pair": flops,
                     "timing": "CUDA events around the force-kernel launch inside the timed step (the step's first kernel "
                               "prefetches the packed list into L2); the second figure is the kernel launched alone right after an L2 flush",
                     "peak_note": "148 SMs x 128 FP32 lanes x 2 x %.0f MHz (clocks.max.sm, MEASURED_PEAKS.json %s)"
                                  % (peaks["sm_max_mhz"], peaks["source"]),
                     "useful_pairs_per_s_kernel": npairs / (k_ms * 1e-3),
                     "computed_pairs_per_s_kernel": ntiles * 64 / (k_ms * 1e-3),
                     "traffic": cap.get("bytes"),
                     "ncu": {k: v for k, v in cap.items() if k != "bytes"} or None,
                     "hbm": {"algorithmic_bytes": int(alg_bytes), "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                             "peak_gbs": peaks["hbm_gbs"], "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}},
        "e2e": {"value": npairs / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(s.n * 12), "d2h_bytes_per_step": int(s.n * 12)},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if full:
        del flush
        if not args.no_search:
            out["search"] = search_block(args, s, coul, local_rank, step_ms, k_ms, h, x_dev)
        if not args.no_sustained:
            # after the re-searches above the context holds a freshly built list of the same coordinates
            out["sustained"] = sustained_block(h, s, npairs, seconds=args.sustained_seconds, flops_per_pair=flops, device=local_rank)
        # ---- CPU baseline: the reference's own SIMD path on this host ----------------------------------------
        cpu = None
        if not args.no_cpu:
            try:
                cores = host_cores()
                t, tk, n = time_reference(s, args.eel, cores, 2000, 3, budget_s=15.0)
                cpu = {"value": npairs / t, "unit": "pairs/s", "cores": cores, "kind": "reference",
                       "sample": "%d full steps of %s, one OpenMP thread on each of the %d physical cores of %s (x convert + 2xMM SIMD kernel + "
                                 "f reduce), %.3f ms/step; kernel alone %.3f ms" % (n, workload, cores, cpu_model(), t * 1e3, tk * 1e3),
                       "search_ms": getattr(time_reference, "last_search_ms", None),
                       "search_note": "the reference's own pair-search step (nbnxn_put_on_grid + constructPairlist) on the same threads"}
            except Exception as e:  # the checker library is optional for the GPU numbers
                cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}
        out["cpu_baseline"] = cpu
    fc.nb.close()
    return out


def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        return run_multi_gpu(args, rank, world, local_rank)
    m = measure_single(args, args.workload, local_rank, full=True)
    out = {"metric": METRIC, "value": m["value"], "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
           "data": "synthetic"}
    out.update({k: m[k] for k in ("config", "details", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks") if k in m})
    for k in ("search", "sustained"):
        if k in m:
            out[k] = m[k]
    if args.secondary and args.secondary != args.workload:
        # BASELINE.json configs[1] beside the default configs[3]: same measurements, fewer steps of bookkeeping
        m2 = measure_single(args, args.secondary, local_rank, full=False)
        out["secondary"] = {"metric": METRIC, "unit": "pairs/s", "value": m2["value"], "ms_per_step": m2["ms_per_step"], "config": m2["config"],
                            "details": m2["details"], "roofline": m2["roofline"], "e2e": m2["e2e"], "gpu_launches": m2["gpu_launches"],
                            "clocks": m2["clocks"]}
    print(json.dumps(out))


def run_multi_gpu(args, rank, world, local_rank):
    """N ranks, one per GPU: slab decomposition along x of a box holding N copies of the N=1 workload (weak scaling) or of
    the workload itself (strong), halos through peer-memory windows (gmxapi_b200/domdec.py)."""
    import torch
    import torch.distributed as dist
    import gmxapi_b200 as g
    from gmxapi_b200 import domdec
    # stdout carries exactly one JSON line: while the job runs, file descriptor 1 points at stderr, so that whatever the libraries
    # underneath print (NCCL's version line under NCCL_DEBUG=VERSION / WARN / INFO) cannot end up in front of it
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = measured_peaks()
    s = workload_system(args.workload, world if args.scaling == "weak" else 1)
    coul = g.CoulombType.Pme if args.eel == "ewald" else g.CoulombType.ReactionField
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=coul, computeVirialAndEnergy=False, device=local_rank, epsilonRf=0.0,
                            maxTilesPerEntry=args.max_tiles)
    if args.dd_grid:
        # 2-D / 3-D decomposition (half-shell rule, halos through NCCL send/recv: gmxapi_b200/domdec_nd.py); --scaling strong only
        from gmxapi_b200 import domdec_nd
        grid = tuple(int(v) for v in args.dd_grid.lower().split("x"))
        if len(grid) != 3 or int(np.prod(grid)) != world:
            raise SystemExit("--dd-grid NXxNYxNZ must multiply to --gpus")
        if args.scaling != "strong":
            raise SystemExit("--dd-grid goes with --scaling strong")
        d = domdec_nd.DomainRankND(s, opt, domdec.TorchDistTransport(), grid, device=local_rank)
    else:
        d = domdec.DomainRank(s, opt, domdec.TorchDistTransport(), device=local_rank)
    h, stream = d.nb, d.stream
    dev = torch.device("cuda", local_rank)
    cnt = torch.tensor([d.pair_count(RC), d.plan.nhome, d.plan.nhalo, h.stats()["ntiles_packed"]], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt)
    npairs, natoms, nhalo_tot, ntiles = (int(v) for v in cnt.tolist())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if not args.no_flush else None

    for _ in range(args.warmup):
        d.step(0)
    h.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = h.stats()["nlaunches"]
    align = torch.zeros(1, dtype=torch.float32, device=dev)
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            if flush is not None:
                flush.fill_(k & 0xff)
            # the 256 MiB flush takes ~43 us +- a few on every rank: without re-aligning the ranks after it, each timed step
            # would also measure how much later the neighbour finished its flush.  Outside the timed region.
            dist.all_reduce(align)
        ev[k][0].record(stream)
        d.step(0)
        ev[k][1].record(stream)
    h.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    launches = h.stats()["nlaunches"] - l0
    t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev]))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    k_ms = h.time_force_kernel(-1, 0, 3, max(10, min(args.steps, 50)), flush_l2=not args.no_flush)
    kt = torch.tensor([k_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    k_ms = float(kt.item())
    clocks = sampler.stop()

    # end to end: pinned host coordinates of the home atoms in, pinned host forces out, every step
    x_pin = torch.from_numpy(np.ascontiguousarray(s.x[d.plan.home])).pin_memory()
    f_pin = torch.empty_like(x_pin).pin_memory()
    for _ in range(args.warmup):
        d.compute(x_pin, 0, f_pin)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d.compute(x_pin, 0, f_pin)
    torch.cuda.synchronize()
    dist.barrier()
    te = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    dist.all_reduce(lt)

    # ---- parity of the path that was just timed (outside every timed region): the decomposed forces, gathered from the ranks'
    # pinned output buffers of the last e2e step, against a single-domain run of the same box on rank 0's GPU ----
    parity = None
    if not args.no_parity:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((np.asarray(d.plan.home), f_pin.numpy().copy()), parts, dst=0)
        if rank == 0:
            f_dd = np.zeros((s.n, 3), np.float32)
            seen = np.zeros(s.n, np.int32)
            for home, fh in parts:
                f_dd[home] = fh
                seen[home] += 1
            assert np.all(seen == 1), "every atom must be home on exactly one rank"
            fc1 = g.ForceCalculator(g.SimulationState.from_system(s), opt)
            f_one = fc1.compute()
            n_one = fc1.nb.pair_count(RC)
            fc1.nb.close()
            rel = float(np.sqrt(((f_dd.astype(np.float64) - f_one) ** 2).sum() / (f_one.astype(np.float64) ** 2).sum()))
            # a pair whose r^2 sits within float rounding of rc^2 AND straddles a periodic edge of a decomposed dimension may flip:
            # the halo carries x_j + box (rounded), the single-domain kernel evaluates (x_i - box) - x_j -- the reference's own DD
            # runs differ from its single-rank runs the same way (domdec.cpp:300-318); tests/test_gpu_domdec.py checks every such pair
            allowed = max(2, int(n_one) // 500000)
            parity = {"force_rel_rms_vs_single_domain": rel, "pairs_decomposed": int(npairs), "pairs_single_domain": int(n_one),
                      "pair_count_tolerance": allowed,
                      "checked": "forces of the last end-to-end step of every rank, gathered; pair counts summed over ranks (pairs at "
                                 "r^2 = rc^2 +- rounding across a periodic edge may flip, see tests/test_gpu_domdec.py)"}
            if not (rel <= 1e-5 and abs(n_one - npairs) <= allowed):
                raise SystemExit("decomposed run disagrees with the single-domain run: %r" % (parity,))
    if rank == 0:
        flops = FLOPS_PER_PAIR[args.eel]
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12 * world
        achieved = npairs * flops / (k_ms * 1e-3) / 1e12
        halo_bytes = nhalo_tot * 12
        out = {
            "metric": METRIC, "value": npairs / (step_ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(args.workload, s, npairs, args.eel, args.scaling, world),
            "details": {"computed_pairs_per_step": int(ntiles * 64),
                        "decomposition": ("%s over %s domains, half shell" % (args.workload, args.dd_grid)) if args.dd_grid
                        else "%d x-slabs" % world,
                        "parallelism": ("dd%s" % args.dd_grid) if args.dd_grid else "dd%dx1x1" % world,
                        "halo": "peer-memory windows over NVLink, written and awaited inside the step's kernels",
                        "l2": "L2 flushed (256 MiB write) between timed steps" if not args.no_flush else "not flushed",
                        "halo_atoms_total": int(nhalo_tot), "halo_bytes_per_step_each_way": int(halo_bytes)},
            "parity": parity,
            "roofline": {"bound": "fp32", "kernel": "k_force (local + non-local), slowest rank", "achieved": achieved, "peak": fp32_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak, "kernel_ms": k_ms, "flops_per_useful_pair": flops,
                         "peak_note": "%d GPUs x 148 SMs x 128 FP32 lanes x 2 x %.0f MHz" % (world, peaks["sm_max_mhz"]), "traffic": None},
            "cpu_baseline": None,
            "e2e": {"value": npairs / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(natoms * 12), "d2h_bytes_per_step": int(natoms * 12)},
            "gpu_launches": int(lt.item()),
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    d.close()
    del flush, x_pin, f_pin
    torch.cuda.synchronize()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="water_1M", help="water_24k / water_96k / water_192k / water_1M / water_1.5M (jittered lattice) or "
                    "ref_water_24k / ref_water_96k / ref_water_192k / ref_water_1M (the reference's benchmark water)")
    ap.add_argument("--secondary", default="water_24k", help="N = 1: a second workload reported in the `secondary` block ('' = none)")
    ap.add_argument("--eel", default="ewald", choices=["ewald", "rf"])
    ap.add_argument("--max-tiles", type=int, default=0, help="cluster pairs per list entry (0 = library default)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the workload itself over N domains (default), weak = N copies of the workload along x")
    ap.add_argument("--dd-grid", default="", help="N > 1: decompose as NXxNYxNZ ranks (e.g. 2x2x2) instead of x slabs")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-search", action="store_true", help="skip the search / prune / amortised block")
    ap.add_argument("--no-sustained", action="store_true", help="skip the sustained-clock block")
    ap.add_argument("--sustained-seconds", type=float, default=2.5)
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the check against a single-domain run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
