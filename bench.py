#!/usr/bin/env python
"""bench.py — particle-steps/s of the SPH step loop (BASELINE.json metric) on N B200s, plus the CPU arm.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA library through the C ABI)
  python bench.py --impl reference [...]                       the reference arm: the CPU restatement (oracle/,
                                                               kind "port" — the Rust reference cannot be built
                                                               in this image) on all host cores

Workload (every N): the north star's case — 2-D dam break, 15 996 516 initial particles (`default-scene-web.yaml`
block at spacing 2.806e-4), adaptive h with radius ratio 4:1, HybridDFSPH, EmptyAngle level set and share / merge /
split every step (SURVEY.md §8d C3 recipe: particle_radius_fine = the radius of a lattice particle,
particle_radius_base = 4x, maximum_surface_distance 0.2, default-config otherwise).  The scene is advanced by
PREROLL_STEPS untimed steps first (input preparation): the interior coarsens towards the base size, the particle
count settles near 5 M, the block lands.  During the first RAMP_STEPS of them particle_radius_base grows linearly
from the fine radius to its final value — parameters are passed by value every step, as in the reference
(main_loop.rs:280) — because merging two thirds of 16 M particles in ONE step, which the plain recipe does, throws
thousands of surface particles off at 10-20 m/s in the reference algorithm (CPU oracle and GPU alike) and the CFL
rule then holds the whole simulation at dt ~ 1e-5.
A "step" is one FluidSimulation::single_step (physics + resampling) over the whole particle set.
value = sum over the timed steps of N_step / device time of those steps (CUDA events on the library stream).
At N > 1 the SAME scene is cut into N x-slabs (strong scaling): bench_dist.py.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec on 2D dam-break at 1/2/4/8 B200; HBM GB/s vs roofline"
UNIT = "particle-steps/s"
SPACING_1M = 1.122e-3     # BASELINE configs[1]:   623 x 1604 =    999 292 particles
SPACING_4M = 5.612e-4     # BASELINE configs[2]:  1247 x 3207 =  3 999 129
SPACING_16M = 2.806e-4    # north star:           2494 x 6414 = 15 996 516
SPACING_C2 = SPACING_1M   # (name used by tests / tools for the configs[1] spacing)
SPACING = float(os.environ.get("ASPH_BENCH_SPACING", SPACING_16M))  # smaller scenes for dry runs of this script
RATIO = 4.0
DMAX = 0.2
# The block of the reference's dam-break scene (default-scene-web.yaml) starts 0.1 above the floor and hits it at
# 1.4 m/s, an impact the reference's Jacobi solver does not survive reliably at a million particles and more (CPU oracle
# and GPU alike: DESIGN.md §5); the benchmark lowers the drop to DROP_GAP (same block, spacing, fill ratio).
DROP_GAP = 0.02
PREROLL_T = 0.064         # uniform configs[1] line: simulated time the scene is advanced to (ten steps after the landing)
PREROLL_STEPS = int(os.environ.get("ASPH_BENCH_PREROLL", 120))
RAMP_STEPS = 60
BLOCK_W, BLOCK_H = 0.7, 1.8

# algorithmic bytes per particle and launch of the kernels that are timed one by one
KERNEL_BYTES = {
    "neighbors": (20.0, "k_neighbors (K2+K7+K9+K11: neighbour lists + density + a_ii partials; x,h,m -> rho,a_ii: 20 B/particle algorithmic)"),
    "accel_sweep": (28.0, "k_sweep<0> (K14, pressure-acceleration pass: x,m,rho,p -> a^p; 28 B/particle algorithmic)"),
    "jacobi_sweep": (40.0, "k_sweep<1> (K15, Jacobi update pass: x,m,rho,a^p,p,s,a_ii -> p'; 40 B/particle algorithmic)"),
    "level_propagate": (20.0, "k_propagate (K4, level-set propagation, one persistent launch per step; 20 B/particle algorithmic in total)"),
    "partner_search": (32.0, "k_greedy (K19/K20 partner search, one persistent launch per share / merge phase; x,m,level,class in, partner out: 32 B/particle)"),
}


def default_params(A):
    return A.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml"))


def uniform_params(A):
    """The reference's "Uniform SPH" recipe (media/motivation-video.yaml:42-57): BASELINE configs[1]."""
    return default_params(A).replace(merging=False, sharing=False, splitting=False, level_estimation_method="None")


def fine_radius(spacing):
    return math.sqrt(0.93 / math.pi) * spacing  # radius of a lattice particle: volume_fill_ratio * spacing^2 = pi r^2


def adaptive_params(A, spacing, ratio=RATIO):
    r_f = fine_radius(spacing)
    return default_params(A).replace(particle_radius_fine=r_f, particle_radius_base=ratio * r_f, maximum_surface_distance=DMAX)


def ramped(params, spacing, step, ratio=RATIO, ramp_steps=RAMP_STEPS):
    """Parameters of pre-roll step number `step`: the base radius on its way from the fine radius to ratio x fine."""
    f = min(1.0, (step + 1) / float(ramp_steps))
    return params.replace(particle_radius_base=(1.0 + (ratio - 1.0) * f) * fine_radius(spacing))


def dam_break(A, spacing, n_gpus=1, block_width=None, kind=None):
    """The dam-break block DROP_GAP above the floor.  (n_gpus / kind: the widened tanks of the round-1 weak-scaling runs,
    kept for tools/ and the tests of their geometry; the benchmark itself uses the one 2 m tank at every N.)"""
    w = 2.0 * n_gpus
    bw = BLOCK_W if block_width is None else block_width
    y0 = -1.0 + DROP_GAP
    if n_gpus == 1 or (kind or "wide") == "wide":
        return A.SceneConfig.dam_break(spacing, pos=(-w / 2 + 0.05, y0), size=(bw * n_gpus, BLOCK_H), width=w, height=2.0)
    spans = [(-w / 2 + 0.05, bw / 2)] + [(-w / 2 + 2.0 * k - bw / 2, bw) for k in range(1, n_gpus)] + [(w / 2 - 0.05 - bw / 2, bw / 2)]
    return A.SceneConfig({"boundary": {"type": "box", "width": w, "height": 2.0},
                          "blocks": [{"pos": [x0, y0], "size": [width, BLOCK_H], "spacing": spacing, "volume_fill_ratio": 0.93,
                                      "velocity": [0, 0]} for x0, width in spans]})


def workload_name(n0):
    return (f"north star: 2D dam-break, {n0} initial particles (default-scene-web block at spacing {SPACING:g}, {DROP_GAP} above the floor), adaptive h "
            f"radius ratio {RATIO:g}:1 (particle_radius_fine = lattice particle radius, maximum_surface_distance {DMAX}), HybridDFSPH, EmptyAngle level set, "
            f"share / merge / split every step; state after {PREROLL_STEPS} untimed steps (base radius ramped up over the first {RAMP_STEPS})")


def preroll(sim, t_target, max_steps=2000):
    """Advance a scene to simulated time t_target (input preparation, untimed).  Returns the steps taken."""
    k = 0
    while sim.time < t_target and k < max_steps:
        sim.single_step()
        k += 1
    return k


def preroll_adaptive(sim, A, base, spacing, steps=None):
    """The adaptive scene's input preparation: `steps` untimed steps, base radius ramped up over the first RAMP_STEPS."""
    steps = PREROLL_STEPS if steps is None else steps
    for s in range(steps):
        sim.single_step(ramped(base, spacing, s))
    return steps


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def issue_roofline(n, launch_ms, sm_count, sm_mhz, warp_instructions, hbm_peak_gbs):
    """Second reading of a pair pass: its ceiling is the issue rate (4 warp instructions / clock / SM), reported
    beside the contract's HBM roofline.  `warp_instructions` per launch comes from a committed ncu capture (profiles/)."""
    peak = 4.0 * sm_count * sm_mhz * 1e6
    achieved = warp_instructions / (launch_ms * 1e-3)
    floor_s = warp_instructions / peak                      # duration with every issue slot filled
    return {"bound": "issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s", "frac": achieved / peak,
            "sm_count": int(sm_count), "sm_mhz": float(sm_mhz), "instructions_per_launch": float(warp_instructions),
            "instructions_per_particle": 32.0 * warp_instructions / n,
            # the contract's HBM fraction (40 B x particles / time / peak) this instruction mix would reach at full issue
            "hbm_frac_at_full_issue": 40.0 * n / floor_s / 1e9 / hbm_peak_gbs}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown," \
        "clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap"

    def __init__(self, device=0):
        self.device, self.rows, self.proc, self.thread = device, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def pinned(shape, dtype=np.float32):
    """numpy view of page-locked host memory (cudaHostAlloc through torch)."""
    import torch
    t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    return t.numpy(), t


class StepLog:
    """What a run of steps did: the figures throughput is meaningless without."""
    KEYS = ("div_sweeps", "density_sweeps", "level_sweeps", "n_shared", "n_merged", "n_split_parents")

    def __init__(self):
        self.n, self.rows, self.rounds, self.dt = [], [], [], []

    def add(self, sim):
        i = sim.step_info()
        self.n.append(int(i["n_particles_begin"]))
        self.rows.append([int(i[k]) for k in self.KEYS])
        self.rounds.append(sim.adapt_rounds())
        self.dt.append(float(i["dt"]))

    def summary(self):
        a = np.asarray(self.rows, dtype=np.float64).reshape(-1, len(self.KEYS))
        out = {"avg_" + k: float(a[:, j].mean()) for j, k in enumerate(self.KEYS)} if len(a) else {}
        out.update({"particles_first": self.n[0] if self.n else None, "particles_last": self.n[-1] if self.n else None,
                    "avg_greedy_rounds": float(np.mean(self.rounds)) if self.rounds else None,
                    "max_greedy_rounds": int(max(self.rounds)) if self.rounds else None,
                    "avg_dt": float(np.mean(self.dt)) if self.dt else None})
        return out


def kernel_table(kt, K, n_avg, peak):
    """Per timed kernel class: average launch, launches and milliseconds per step, algorithmic GB/s and roofline fraction."""
    tab = {}
    for name, (ms, cnt) in kt.items():
        if cnt == 0:
            continue
        row = {"avg_launch_ms": ms / cnt, "launches_per_step": cnt / max(K, 1), "ms_per_step": ms / max(K, 1)}
        if name in KERNEL_BYTES:
            gbs = KERNEL_BYTES[name][0] * n_avg / (ms / cnt * 1e-3) / 1e9
            row.update({"algorithmic_bytes_per_particle": KERNEL_BYTES[name][0], "achieved_gbs": gbs, "frac": gbs / peak})
        tab[name] = row
    return tab


def run_ours(args):
    import asph_b200 as A
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        from bench_dist import run  # multi-GPU arm: the same scene in N x-slabs
        return run(args, A, rank, world)
    t_start = time.perf_counter()
    lib = A.load_library()
    base = adaptive_params(A, SPACING)
    scene = dam_break(A, SPACING)
    base = A.init_simulation_params(base, scene)
    split = A.load_split_patterns_from_file()
    sim = A.init_fluid_sim(base, scene, split, counters_enabled=True, lib=lib)
    assert sim.backend() == "cuda-sm100a"
    boundary = A.scene_boundary(scene, base["init_boundary_handler"])
    n0 = sim.num_fluid_particles()
    K, W = args.steps, args.warmup

    # ---- input preparation (untimed), warm-up ------------------------------------------------------------------
    t_pre = time.perf_counter()
    preroll_adaptive(sim, A, base, SPACING)
    t_pre = time.perf_counter() - t_pre
    for _ in range(W):
        sim.single_step(base)
    warm_state = (sim.get_field("position"), sim.get_field("velocity"), sim.get_field("mass"))
    warm_step_number = sim.step_number
    print(f"[bench] input prepared: {n0} -> {len(warm_state[2])} particles in {t_pre:.1f} s; {sim.kernel_launches()} kernel launches before the timed region",
          file=sys.stderr, flush=True)

    # ---- device-resident arm: K timed steps; time = CUDA events around every step on the library stream ---------
    sim.set_kernel_timing(1)
    cnt0 = sim.counters()
    l0 = sim.kernel_launches()
    clocks = ClockSampler()
    clocks.start()
    t0 = time.perf_counter()
    log = StepLog()
    dev_ms = 0.0
    for _ in range(K):
        c_before = sim.counters()["simulation-step"][0]
        sim.single_step(base)
        dev_ms += sim.counters()["simulation-step"][0] - c_before
        log.add(sim)
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    cnt1 = sim.counters()
    phases = {k: (cnt1[k][0] - cnt0[k][0]) / max(K, 1) for k in cnt1}
    launches = sim.kernel_launches() - l0
    kt = sim.kernel_timing()
    sim.set_kernel_timing(0)
    particle_steps = float(sum(log.n))
    n_avg = particle_steps / max(K, 1)
    value = particle_steps / (dev_ms * 1e-3)
    sm = log.summary()

    # ---- rooflines: the dominant kernel (largest share of the step among the kernels timed one by one), the step ----
    peak, peak_src = peaks()
    tab = kernel_table(kt, K, n_avg, peak)
    single = {k: v for k, v in tab.items() if k in KERNEL_BYTES}
    dom = max(single, key=lambda k: single[k]["ms_per_step"]) if single else None
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
    if dom:
        roof.update({"kernel": KERNEL_BYTES[dom][1], "achieved": single[dom]["achieved_gbs"], "frac": single[dom]["frac"],
                     "avg_launch_ms": single[dom]["avg_launch_ms"], "share_of_step": single[dom]["ms_per_step"] / (dev_ms / max(K, 1)),
                     "particles_per_launch": n_avg})
        try:  # DRAM bytes of one launch of that kernel from the committed ncu capture of this workload (profiles/)
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tr = json.load(f)[dom]
            roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof["traffic_source"] = tr.get("source", "profiles/r2_traffic.json (ncu --set full, per launch)")
        except (OSError, KeyError, ValueError):
            pass
    s_div, s_den = sm.get("avg_div_sweeps", 0.0), sm.get("avg_density_sweeps", 0.0)
    b_step = 388.0 + 68.0 * (s_div + s_den)  # adaptive DFSPH, B = 388 + 68 (S_div + S_den) bytes per particle-step
    step_roof = {"bytes_per_particle_step": b_step, "achieved": b_step * value / 1e9, "frac": b_step * value / 1e9 / peak, "unit": "GB/s"}

    # ---- e2e arm: host buffers in, host buffers out, every step (same K steps from the same warm state) ---------
    cap = int(len(warm_state[2]) * 1.1) + 4096
    hp, _tp = pinned((cap, 2)); hv, _tv = pinned((cap, 2)); hm, _tm = pinned((cap,))

    def load_warm():
        n = len(warm_state[2])
        hp[:n] = warm_state[0]; hv[:n] = warm_state[1]; hm[:n] = warm_state[2]
        sim.lib.asph_set_step_number(sim._h, warm_step_number)
        return n

    n_cur = load_warm()
    sim.set_state(hp[:n_cur], hv[:n_cur], hm[:n_cur]); sim.single_step(base)  # one untimed pass through this path
    n_cur = load_warm()
    e2e_log = StepLog()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for _ in range(K):
        sim.set_state(hp[:n_cur], hv[:n_cur], hm[:n_cur])    # H2D of this step's inputs (x, v, m) from pinned host memory
        sim.single_step(base)
        e2e_log.add(sim)
        h2d += n_cur * 20
        n_cur = sim.num_fluid_particles()                      # resampling changes the particle count
        sim.get_field("position", out=hp[:n_cur])              # D2H of the step's result (x, v, m), reference particle order,
        sim.get_field("velocity", out=hv[:n_cur])              # into the same host buffers: this step's output is the next
        sim.get_field("mass", out=hm[:n_cur])                  # step's input
        d2h += n_cur * 20
    e2e_s = time.perf_counter() - t0
    e2e = {"value": float(sum(e2e_log.n)) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(K, 1)), "d2h_bytes_per_step": int(d2h / max(K, 1)),
           "ms_per_step": e2e_s * 1e3 / max(K, 1), "work": e2e_log.summary()}
    same_work = all(abs(e2e["work"].get(k, 0) - sm.get(k, 0)) <= 1.0 for k in ("avg_div_sweeps", "avg_density_sweeps", "avg_level_sweeps"))
    e2e["same_work_as_device_arm"] = bool(same_work)
    sim_time = sim.time
    dup = sim.greedy_duplicates()
    sim.close()
    if not same_work:
        raise SystemExit(f"bench: the e2e arm did not do the work of the device arm: {e2e['work']} vs {sm}")

    # ---- CPU baseline: the oracle (port of the reference) from the same warm state, bounded sample -----------------
    cpu = cpu_baseline_from(A, base, boundary, split, warm_state, warm_step_number, budget_s=args.cpu_budget) if args.cpu_budget > 0 else None

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / max(K, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict({"workload": workload_name(n0), "particles_initial": n0, "preroll_steps": PREROLL_STEPS, "preroll_wall_s": t_pre,
                        "l2": "working set per step (neighbour lists + SoA, > 1 GB) exceeds the 126 MB L2; no flush",
                        "timing": "CUDA events on the library stream around every step (PerformanceCounters 'simulation-step': physics + resampling)",
                        "level_sweeps_note": "the propagation stops once a whole sweep has assigned only values below -maximum_surface_distance (they are clamped downstream, simulation.rs:833-836: identical results); the reference and the CPU arms sweep on until nothing changes (about 230 sweeps on this state)",
                        "wall_ms_per_step": wall * 1e3 / max(K, 1), "phase_ms_per_step": phases, "simulated_time": sim_time, "greedy_duplicates": dup,
                        "kernels": tab, "switches": {k: os.environ[k] for k in ("ASPH_BULK", "ASPH_SWEEP_GRID", "ASPH_BENCH_SPACING", "ASPH_BENCH_PREROLL") if k in os.environ}}, **sm),
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "step_roofline": step_roof, "cpu_baseline": cpu,
    }
    if not args.no_secondary:
        try:
            out["secondary"] = {"configs1_uniform_1m": uniform_line(A, lib, peak)}
        except Exception as e:  # informational: never at the cost of the line
            out["secondary"] = {"configs1_uniform_1m": {"error": repr(e)[:300]}}
    out["bench_wall_s"] = time.perf_counter() - t_start
    print(json.dumps(out), flush=True)


def uniform_line(A, lib, peak, steps=12, warmup=3):
    """BASELINE configs[1] (2-D dam break, uniform h, 999 292 particles, HybridDFSPH; the headline of round 1) beside the
    headline: the regime in which both pressure solves iterate, i.e. the sweep kernels' own numbers."""
    params = uniform_params(A)
    scene = dam_break(A, SPACING_1M)
    pos, vel, mass = A.scene_particles(scene)
    n = len(mass)
    sim = A.FluidSimulation(params, pos, vel, mass, A.scene_boundary(scene, "AnalyticOverestimate"), counters_enabled=True, lib=lib)
    pre = preroll(sim, PREROLL_T)
    for _ in range(warmup):
        sim.single_step()
    sim.set_kernel_timing(1)
    dev_ms, log = 0.0, StepLog()
    for _ in range(steps):
        c0 = sim.counters()["simulation-step"][0]
        sim.single_step()
        dev_ms += sim.counters()["simulation-step"][0] - c0
        log.add(sim)
    kt = sim.kernel_timing()
    sim.close()
    sm = log.summary()
    value = n * steps / (dev_ms * 1e-3)
    b_step = 260.0 + 68.0 * (sm["avg_div_sweeps"] + sm["avg_density_sweeps"])
    return {"workload": f"configs[1]: 2D dam-break, uniform h, {n} particles, HybridDFSPH, state at t = {PREROLL_T} s ({pre} untimed steps)",
            "value": value, "unit": UNIT, "ms_per_step": dev_ms / steps, "steps": steps, "avg_div_sweeps": sm["avg_div_sweeps"],
            "avg_density_sweeps": sm["avg_density_sweeps"], "particle_sweeps_per_s": value * (sm["avg_div_sweeps"] + sm["avg_density_sweeps"]),
            "kernels": kernel_table(kt, steps, n, peak),
            "step_roofline": {"bytes_per_particle_step": b_step, "achieved": b_step * value / 1e9, "frac": b_step * value / 1e9 / peak}}


def load_oracle(A):
    oracle_path = os.path.join(ROOT, "oracle", "liboracle_f32.so")
    if not os.path.exists(oracle_path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j", "2"], stdout=subprocess.DEVNULL)
    olib = A.load_library(oracle_path)
    olib.oracle_max_threads.restype = C.c_int
    return olib, int(olib.oracle_max_threads())


def cpu_baseline_from(A, params, boundary, split, state, step_number, budget_s, max_steps=200):
    """The CPU oracle stepping the SAME state with the same parameters: whole steps until the time budget is used."""
    olib, cores = load_oracle(A)
    o = A.FluidSimulation(params, state[0], state[1], state[2], boundary, split, lib=olib)
    o.lib.asph_set_step_number(o._h, step_number)
    t0 = time.perf_counter()
    log = StepLog()
    while len(log.n) < 1 or ((time.perf_counter() - t0) < budget_s and len(log.n) < max_steps):
        o.single_step(params)
        log.add(o)
    el = time.perf_counter() - t0
    o.close()
    sm = log.summary()
    return {"value": float(sum(log.n)) / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(log.n)} whole steps of the same {len(state[2])}-particle state (post-warm-up), same parameters: {el:.1f} s of CPU time, "
                      f"OpenMP over particles, serial phases as in the reference; sweeps div {sm['avg_div_sweeps']:.1f} density {sm['avg_density_sweeps']:.1f} "
                      f"level {sm['avg_level_sweeps']:.0f}",
            "work": sm}


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path on the SAME workload.  The Rust crate cannot be
    built here (no cargo/rustc), so this is the oracle port (kind 'port') with all host threads.  Its input state — the
    scene after the untimed pre-roll and the warm-up steps — is prepared exactly as in our arm (input preparation runs on
    the GPU library when a GPU is present: 120 steps of 16 M particles take the CPU hours; nothing of it is timed);
    the timed steps are whole steps of that state, as many of the K asked for as fit --ref-budget seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all host threads the process may use: torchrun pins OMP_NUM_THREADS to 1 for its workers, which would turn the
    # reference arm into a single-core run (the OpenMP runtime reads the variable when the oracle library is loaded)
    try:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    except AttributeError:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import asph_b200 as A
    olib, cores = load_oracle(A)
    split = A.load_split_patterns_from_file()
    spacing, prepared_by = SPACING, None
    try:
        lib = A.load_library()
        base = adaptive_params(A, spacing)
        scene = dam_break(A, spacing)
        base = A.init_simulation_params(base, scene)
        sim = A.init_fluid_sim(base, scene, split, lib=lib)
        prepared_by = "the CUDA library (untimed input preparation, as in our arm)"
    except Exception as e:  # no GPU: a scene the CPU can prepare itself in minutes (the configs[1] spacing)
        spacing = max(SPACING, SPACING_1M)
        base = adaptive_params(A, spacing)
        scene = dam_break(A, spacing)
        base = A.init_simulation_params(base, scene)
        sim = A.init_fluid_sim(base, scene, split, lib=olib)
        prepared_by = f"the CPU oracle itself at spacing {spacing:g} (no usable GPU for the full-size preparation: {str(e)[:80]})"
    n0 = sim.num_fluid_particles()
    t_pre = time.perf_counter()
    preroll_adaptive(sim, A, base, spacing)
    for _ in range(args.warmup):
        sim.single_step(base)
    t_pre = time.perf_counter() - t_pre
    state = (sim.get_field("position"), sim.get_field("velocity"), sim.get_field("mass"))
    step_number = sim.step_number
    boundary = A.scene_boundary(scene, base["init_boundary_handler"])
    sim.close()
    o = A.FluidSimulation(base, state[0], state[1], state[2], boundary, split, lib=olib)
    o.lib.asph_set_step_number(o._h, step_number)
    log = StepLog()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.single_step(base)
        log.add(o)
        if time.perf_counter() - t0 > args.ref_budget:
            break
    el = time.perf_counter() - t0
    o.close()
    done = len(log.n)
    value = float(sum(log.n)) / el
    sm = log.summary()
    same = spacing == SPACING
    sample = (f"{done} whole steps (of {args.steps} asked) of the {len(state[2])}-particle state the workload is in after {PREROLL_STEPS} pre-roll + {args.warmup} warm-up steps "
              f"({n0} initial particles; prepared in {t_pre:.0f} s by {prepared_by}); all host threads, OpenMP over particles, serial phases as in the reference; "
              f"sweeps div {sm['avg_div_sweeps']:.1f} density {sm['avg_density_sweeps']:.1f} level {sm['avg_level_sweeps']:.0f}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": el * 1e3 / max(done, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict({"workload": workload_name(n0) if same else f"the same recipe at spacing {spacing:g} ({n0} initial particles): CPU-only box",
                        "particles_initial": n0, "preroll_steps": PREROLL_STEPS, "same_config_as_our_arm": same}, **sm),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU time for the cpu_baseline leg (0 = skip)")
    ap.add_argument("--ref-budget", type=float, default=120.0, help="time cap of the reference arm's timed steps")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] (uniform h, 1 M particles) line reported under \"secondary\"")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
