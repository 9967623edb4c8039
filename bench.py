#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the explicit MPM step (3D uGIMP, isotropic elastic, FLIP, USAVG+).

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path (default arm)
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's own CPU path (oracle/_ref)

One "step" = one full MPMStep (tasks 1-9, 11) over the resident particle block.  Workloads
(BASELINE.json configs, synthetic lattice blocks exactly as the reference's generator makes them):
  block8m  (default)  100^3 cells x 8 particles = 8,000,000 particles per GPU: config 5 at N GPUs
                      (and the size the north_star target is quoted on at N=1)
  block1m             50^3 cells = 1,000,000 particles: config 2
  neo8m               100^3-cell Neo-Hookean block, lCPDI shape functions, XPIC(2) every step, gravity, clamped bottom plane
                      (config 3 family; general per-task kernels, one GPU); neo1m = 50^3
  taylor16m           100x100x200-cell IsoPlasticity bar hitting a plate of rigid-BC particles at 200 m/s,
                      16,000,000 particles TOTAL split over the GPUs (config 4, strong scaling); taylor2m = 50x50x100
Prints ONE JSON line (see README / DESIGN.md for the keys).
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

METRIC = "particle-steps/sec (3D uGIMP) at 1/2/4/8 B200; % of HBM roofline"
UNIT = "particle-steps/s"
ALGO_BYTES_PER_PARTICLE_STEP = 504.0        # elastic uGIMP 3D FLIP USAVG+

# algorithmic (compulsory) bytes per particle of each task's kernels, 3D uGIMP elastic, FP64 SoA
# (DESIGN.md "kernels and their algorithmic bytes"): fields each kernel must read + write once.
TASK_ALGO_BYTES = {
    "initialization": 3 * 8 + 4 + 3 * 8,                       # pos, elem -> ncpos
    "mass_and_momentum": (3 + 3 + 3 + 1) * 8 + 4 + 5,          # ncpos, lp, vel, mp, elem; grid 4.5 B/particle
    "post_extrapolation": 6,                                   # node sweep: 2x3 doubles per node / 8 ppc
    "update_strains_first": (3 + 3) * 8 + 4 + 2 * (9 + 6 + 4) * 8 + 3,   # ncpos, lp, elem; F, sp, energies r+w; grid vk
    "grid_forces": (3 + 3 + 1 + 6 + 1) * 8 + 4 + 3,
    "post_forces": 6,
    "update_momenta": 6,
    "update_particles": (3 + 3) * 8 + 4 + 2 * (3 + 3) * 8 + 3 * 8 + 6,   # ncpos, lp; pos, vel r+w; acc w; grid vk+ftot+mass
    "update_strains_last": (3 + 3 + 3 + 1) * 8 + 4 + 2 * (9 + 6 + 4) * 8 + 6,
    "reset_elements": 3 * 8 + 4,
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# fused path: the four particle kernels, elastic state, uniform lp
FUSED_ALGO_BYTES = {"mass_and_momentum": 84, "update_strains_first": 336, "update_particles": 160, "update_strains_last": 356}

TASK_KERNEL = {"mass_and_momentum": "k_f1_mass_momentum", "update_strains_first": "k_f2_strain_forces",
               "update_particles": "k_f3_update_momentum", "update_strains_last": "k_f4_strain_reset"}


def ncu_traffic(task, workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the task's kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/run_ncu_final.sh); None when there is no capture of it."""
    path = os.path.join(ROOT, "profiles", "r2b_dram_traffic.json")
    if not os.path.exists(path) or task not in TASK_KERNEL:
        return None, None
    d = json.load(open(path))
    if d.get("workload") != workload or d.get("particles") != n:
        return None, None
    for k, v in d["dram_bytes_per_launch"].items():
        if k.startswith(TASK_KERNEL[task]):
            return v, "profiles/r2b_dram_traffic.json (%s)" % k
    return None, None


BLOCK_JITTER = 0.4          # hash-jitter amplitude of the block workloads, in cells (positions move by up to +-0.2 cell)


def block_velocity(ncell):
    """Initial velocity field of the block workloads, the same for this repo's arm and the reference arm: 10 m/s
    compression along z with a sinusoidal perturbation (period = one block, so particles cross cell and slab faces:
    ~0.3 % of a cell per step) plus a shear wave in x."""
    L = float(ncell)

    def vel(pos):
        v = np.zeros_like(pos)
        v[2] = -1.0e4 + 2.0e3 * np.sin(2.0 * np.pi * (pos[2] - 7.0) / L)
        v[0] = 1.0e3 * np.sin(2.0 * np.pi * (pos[1] - 7.0) / L)
        return v
    return vel


def make_problem(workload, ncell_override=None, rank=0, world=1):
    from nairn_mpm_fea_b200 import materials as M, problem
    ncell = {"block8m": 100, "block1m": 50, "taylor16m": 100, "taylor2m": 50, "neo8m": 100, "neo1m": 50}[workload]
    if ncell_override:
        ncell = ncell_override
    if workload.startswith("taylor"):
        # config 4: copper-like von Mises bar (E 117 GPa, yield 400 MPa, linear hardening), 200 m/s onto a rigid plate
        u = M.xml_units(E=117.0e3, rho=8.94, yld=400.0, Ep=100.0)
        mat = M.isoplasticity(u["E"], 0.35, u["rho"], u["yld"], u["Ep"])
        ncz = 2 * ncell
        cz = None if world == 1 else ((ncz * rank) // world, (ncz * (rank + 1)) // world)
        pr = problem.block3d(ncell=ncell, margin=7, velocity=(0.0, 0.0, -2.0e5), jitter_amp=0.4, bottom_bc=False, material=mat,
                             ncell_xyz=(ncell, ncell, ncz), cells_z=cz, rigid_wall=dict(set_direction=4, overhang=2))
        return pr, ncell
    vel = block_velocity(ncell)
    if workload.startswith("neo"):
        # config 3: soft Neo-Hookean solid (G 40 MPa, K 200 MPa), lCPDI, XPIC(2), gravity + the block velocity field
        if world != 1:
            raise SystemExit("bench.py: the neo workloads run on the per-task kernels, one GPU")
        u = M.xml_units(G=40.0, K=200.0, rho=1.0)
        mat = M.neohookean(u["G"], u["K"], u["rho"])
        pr = problem.block3d(ncell=ncell, margin=7, velocity_fn=vel, jitter_amp=BLOCK_JITTER, material=mat,
                             shape=problem.LINEAR_CPDI, gravity=(0.0, 0.0, -9.8e6))
        pr.xpic_order = 2
        return pr, ncell

    if world == 1:
        pr = problem.block3d(ncell=ncell, margin=7, velocity_fn=vel, jitter_amp=BLOCK_JITTER)
    else:
        # config 5: one ncell^3 block per GPU stacked along z; this rank generates only its own slab
        pr = problem.block3d(ncell=ncell, margin=7, velocity_fn=vel, jitter_amp=BLOCK_JITTER,
                             ncell_xyz=(ncell, ncell, ncell * world), cells_z=(ncell * rank, ncell * (rank + 1)))
    return pr, ncell


def time_workload(workload, local, steps=20, warmup=3):
    """Device-resident ms/step of another BASELINE config on this GPU (one line of the `configs` key of the default run)."""
    import torch
    from nairn_mpm_fea_b200 import MpmGpu
    if workload == "disks2d":
        # config 1: the 2D plane-strain uGIMP disk impact the parity tests use (reference dump, tests/golden)
        from nairn_mpm_fea_b200.problem import from_reference_dump
        z = np.load(os.path.join(ROOT, "tests", "golden", "disks2d_ugimp_planestrain.npz"), allow_pickle=False)
        prob = from_reference_dump({k: z[k] for k in z.files if not k.startswith("s1/")})
        n = prob.nparticles
    else:
        prob, _ = make_problem(workload)
        n = int(prob.particles["n_nonrigid"])
    sim = MpmGpu(prob, device=local)
    sim.set_poll_interval(16)
    stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local))
    for _ in range(warmup):
        sim.step(1)
    sim.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        sim.step(1)
    e1.record(stream)
    sim.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sim.close()
    return {"particles": n, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": UNIT, "steps": steps}


def host_state_bytes(pt):
    n = 0
    for k, v in pt.items():
        if isinstance(v, np.ndarray):
            n += v.nbytes
    return n


_JSON_LINES = []          # what rank 0 prints at the very end (see main)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nairn_mpm_fea_b200 import MpmGpu, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libmpmgpu has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = read_peaks()

    prob, ncell = make_problem(args.workload, args.ncell, rank, world)
    if args.xpic_order > 0:
        prob.xpic_order, prob.using_fmpm = args.xpic_order, bool(args.fmpm)
    taylor = args.workload.startswith("taylor")
    neo = args.workload.startswith("neo")
    n = int(prob.particles["n_nonrigid"])          # rigid-BC particles (replicated on every rank) are not counted
    if world == 1:
        sim = MpmGpu(prob, device=local, kernel_path=args.kernel_path, sort_interval=args.sort_interval)
        stepper = sim
        stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local))
    else:
        from nairn_mpm_fea_b200.slab import SlabSim, slab_bounds
        bounds = slab_bounds(prob.depth, 8, 8 + (2 * ncell if taylor else ncell * world), world)
        lo, hi = bounds[rank]
        stepper = SlabSim(prob, prob.particles, lo, hi, rank, world, device=local, capacity_factor=1.2, sort_interval=args.sort_interval)
        sim = stepper.sim
        stream = stepper.stream

    def barrier():
        sim.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput: inputs already in HBM -------------------------------------
    if world == 1:
        sim.set_poll_interval(16)       # status word read every 16th step: the host keeps the stream full (errors surface <= 15 steps late)
    for _ in range(args.warmup):
        stepper.step(1)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sim.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        stepper.step(1)
    ev1.record(stream)
    barrier()
    launches = sim.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    if world == 1:
        sim.set_poll_interval(1)
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    t = torch.tensor([n], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total_particles = int(t.item())
    value = total_particles * args.steps / (ms_total * 1e-3)

    # ---- per-task device times (events around every task; serialises, so done OUTSIDE the timed region)
    sim.set_profiling(True)
    psteps = max(2, min(5, args.steps))
    for _ in range(psteps):
        stepper.step(1)
    tt = sim.task_times()
    sim.set_profiling(False)
    task_ms = {k: v[0] / max(1, v[1]) for k, v in tt.items()}
    dom = max(task_ms, key=lambda k: task_ms[k])
    # IsoPlasticity carries eplast(6), pressure, plastic energy and one history double through both strain updates
    full_state = 2 * (6 + 1 + 1 + 1) * 8 if (taylor or neo) else 0
    algo_step = ALGO_BYTES_PER_PARTICLE_STEP + 2 * full_state
    per_particle = FUSED_ALGO_BYTES.get(dom, TASK_ALGO_BYTES[dom]) if args.kernel_path != 1 else TASK_ALGO_BYTES[dom]
    dom_bytes = (per_particle + (full_state if dom in ("update_strains_first", "update_strains_last") else 0)) * n
    dom_gbs = dom_bytes / (task_ms[dom] * 1e-3) / 1e9
    step_gbs = algo_step * n / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom, args.workload, n)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": dom_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src,
                "kernel_ms": task_ms[dom], "kernel_share_of_step": task_ms[dom] / sum(task_ms.values()),
                "whole_step": {"algorithmic_bytes_per_particle_step": algo_step,
                               "achieved": step_gbs, "frac": step_gbs / hbm_peak},
                "task_ms": task_ms}

    # ---- end to end through the public API with HOST buffers ------------------------------------
    # One archive interval of a run, every byte through host memory: upload the particle state an input file defines (positions,
    # velocities, masses, sizes, elements, materials: a fresh body has no stress or strain yet, so those arrays are not sent --
    # mpmgpu_upload_particles zero-fills them on the device) from pinned host memory, run K steps with the per-step BC values
    # going H2D and the status word coming D2H, then write one particle archive: the records (the reference's binary format,
    # <MPMArchiveOrder> of the reference's examples: position, velocity, stress, strain, work and strain energy, element
    # crossings) are packed on the device and only they come back.
    pt = prob.particles
    lean_keys = ("pos", "vel", "mp", "lp", "in_elem", "matnum", "ids")
    pinned = {}
    for k, v in pt.items():
        if isinstance(v, np.ndarray):
            # state arrays travel only when they hold something the device's defaults (zeros; temperature of the last strain
            # update = 1) do not: a hyperelastic body, for one, starts with B = I and J = 1 in its plastic-strain and history slots
            trivial = (not np.any(v[:5]) and np.all(v[5] == 1.0)) if k == "energies" else not np.any(v)
            if k in lean_keys or not trivial:
                pinned[k] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy()
        else:
            pinned[k] = v
    nb = len(prob.bc_value)
    bcv = np.zeros(nb)
    ARCHIVE_ORDER = "iYYYYNNNNNNNYNNNNY"
    rec = sim.archive_record_size(ARCHIVE_ORDER)
    arch = torch.empty(int(rec * sim.num_particles() * (1.3 if world > 1 else 1.0)) + 4096, dtype=torch.uint8).pin_memory().numpy()    # slabs: particle count changes by migration
    barrier()
    t0 = time.perf_counter()
    sim.upload(pinned)
    for _ in range(args.steps):
        sim.update_velocity_bc_values(bcv, prob.bc_active)
        stepper.step(1)
        st = sim.status()
    n_now = sim.num_particles()
    sim.pack_archive(ARCHIVE_ORDER, out=arch)
    ids = sim.download_ids() if world > 1 else None        # slabs: the records come in device order, the ids say whose they are
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    up_bytes = host_state_bytes(pinned)
    down_bytes = rec * n_now + (4 * n_now if world > 1 else 0)
    e2e = {"value": total_particles * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": (up_bytes / args.steps + nb * 12) * world,
           "d2h_bytes_per_step": (down_bytes / args.steps + 32) * world,
           "what": "one archive interval through host memory: upload of the input state (%s; pinned host) "
                   "+ %d steps with per-step BC values H2D and status D2H + one particle archive (%d-byte records '%s' packed on the device) "
                   "D2H, wall clock" % (", ".join(k for k in pinned if isinstance(pinned[k], np.ndarray)), args.steps, rec, ARCHIVE_ORDER)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if taylor else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("%s: 3D lCPDI Neo-Hookean block, %d^3 cells x 8 = %d particles, XPIC(2), gravity, USAVG+, positions hash-jittered "
                                    "(general per-task kernels)" % (args.workload, ncell, n)) if neo else
                                   ("%s: 3D uGIMP IsoPlasticity (von Mises, linear hardening) bar of %dx%dx%d cells x 8 = %d particles in total, "
                                    "200 m/s onto a plate of %d rigid-BC particles, FLIP, USAVG+, positions hash-jittered"
                                    % (args.workload, ncell, ncell, 2 * ncell, total_particles, prob.nparticles - n)) if taylor else
                                   "%s: 3D uGIMP isotropic-elastic block, %d^3 cells x 8 = %d particles per GPU, FLIP, USAVG+, "
                                   "grid %d^3 cells, particle positions hash-jittered +-0.2 cell off the lattice, 10 m/s compression + sinusoidal perturbation "
                                   "(particles cross cell and slab faces)" % (args.workload, ncell, n, prob.horiz),
                       "particles_per_gpu": n, "nodes": prob.nnodes, "l2_policy": "inputs larger than L2 (%.0f MB state)" % (n * 460 / 1e6),
                       "kernel_path": sim_kernel_path_name(args.kernel_path),
                       "particle_update": ("FMPM(%d)" if prob.using_fmpm else "XPIC(%d)") % prob.xpic_order if prob.xpic_order > 0 else "FLIP",
                       "parallelism": "1 GPU" if world == 1 else
                       "%d z-slabs, one process per GPU: 3 halo-plane exchanges per step + particle migration over NCCL "
                       "(rank 0 sent %d and received %d particle rows)" % (world, stepper.migrated_out, stepper.migrated_in)},
            "roofline": roofline, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches)}
    sim.close()
    if not args.no_cpu_baseline:
        # the reference's CPU path on a bounded sample of the workload (rank 0, host cores), and the parity of this arm
        # against the state that sample ends in (every rank takes part: at N > 1 the sample is run as N slabs)
        import tempfile
        ref_npz = os.path.join(tempfile.mkdtemp(prefix="mpmbench_ref_"), "ref_state.npz") if rank == 0 else None
        if rank == 0:
            line["cpu_baseline"] = cpu_baseline(args.cpu_ncell, args.cpu_steps, dump=ref_npz)
        if world > 1:
            dist.barrier()
        par = parity_leg(args, rank, world, local, ref_npz)
        if rank == 0:
            line["parity"] = par
            line["parity_max_rel"] = par["max_rel"] if par else None
        if world == 1 and args.workload == "block8m" and not args.no_other_configs:
            # the other BASELINE.json configs on this GPU, device-resident (their own lines: profiles/bench_r2/)
            line["configs"] = {"1: 2D plane-strain uGIMP disk impact (tests/golden/disks2d_ugimp_planestrain)": time_workload("disks2d", local, 200, 20),
                               "2: block1m (3D uGIMP elastic block, 1M particles, FLIP)": time_workload("block1m", local, 50, 5),
                               "3: neo8m (3D Neo-Hookean block, 8M particles, lCPDI, XPIC(2))": time_workload("neo8m", local, 10, 3),
                               "4: taylor16m on ONE GPU (IsoPlasticity bar on rigid-BC plate, 16M particles)": time_workload("taylor16m", local, 10, 3),
                               "5: block8m": "this line"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _JSON_LINES.append(json.dumps(line))


def sim_kernel_path_name(k):
    return {0: "auto", 1: "per-task kernels (global atomics)", 2: "tiled"}[k]


_REF_WORKER = r"""
import sys, time, json
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefRun, _flatten
from tests.inputs import block3d
from nairn_mpm_fea_b200.problem import jitter
import bench
import tempfile, os
d = tempfile.mkdtemp(prefix="mpmbench_")
xml = os.path.join(d, "in.fmcmd")
open(xml, "w").write(block3d(ncell=%(ncell)d, margin=7, maxtime=1.0))
os.chdir(d)
t0 = time.perf_counter()
r = RefRun(xml, nprocs=%(nprocs)d)
# the same off-lattice start and velocity field the GPU arm gets (problem.block3d(jitter_amp, velocity_fn))
pos = jitter(r.particles()["pos"], bench.BLOCK_JITTER, 12345)
bad = r.set_particles(pos, bench.block_velocity(%(ncell)d)(pos))
assert bad == 0, bad
dump = %(dump)r
out = {}
def snap(tag, nodes):
    if dump:
        _flatten(tag, r.particles(), out)
        if nodes:
            _flatten(tag + "n", r.nodes(), out)
t1 = time.perf_counter()
r.step(%(warm)d)
t2 = time.perf_counter()
snap("a", %(nodes)d)
t2b = time.perf_counter()
r.step(%(steps)d)
t3 = time.perf_counter()
snap("b", 0)
if dump:
    np.savez(dump, **out)
print(json.dumps({"n": r.info["nmpms"], "setup_s": t1 - t0, "warm_s": t2 - t1, "run_s": t3 - t2b, "patches": r.info["numPatches"]}))
"""


def time_reference(ncell, steps, warm=1, nprocs=None, dump=None, nodes=False):
    """Time the UNMODIFIED reference (oracle/_ref/libnairnmpm_ref.so) on the host cores.  dump: path of an .npz that
    receives the reference's particle state after the warm-up steps (keys a/...) and at the end (b/...); nodes: also the
    node state after the warm-up steps (an/...)."""
    from oracle import refharness
    if not refharness.available():
        return None
    nprocs = nprocs or os.cpu_count() or 1
    code = _REF_WORKER % dict(root=ROOT, ncell=ncell, nprocs=nprocs, warm=warm, steps=steps, dump=dump, nodes=1 if nodes else 0)
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(nprocs)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    if p.returncode != 0:
        return {"error": (p.stderr or p.stdout)[-400:]}
    d = json.loads(p.stdout.strip().splitlines()[-1])
    d["cores"] = nprocs
    d["steps"] = steps
    d["warm"] = warm
    return d


PARITY_FIELDS = ("pos", "vel", "sp", "ep", "wrot")


def parity_leg(args, rank, world, local, ref_npz):
    """Parity of the bench workload itself: the CPU sample of the reference arm (the same block at cpu_ncell^3 cells, same
    jittered start and velocity field) is run by THIS arm too -- on one GPU, or cut into `world` z-slabs with halo
    exchange and particle migration over NCCL -- for the same number of steps; rank 0 compares the final particle state
    with the reference's.  Returns max over fields of max|ours - ref| / max|ref| (None on the other ranks)."""
    import torch.distributed as dist
    from nairn_mpm_fea_b200 import MpmGpu, problem
    ncell = args.cpu_ncell
    nsteps = 1 + args.cpu_steps
    prob = problem.block3d(ncell=ncell, margin=7, velocity_fn=block_velocity(ncell), jitter_amp=BLOCK_JITTER)
    n = prob.nparticles
    if world == 1:
        sim = MpmGpu(prob, device=local, kernel_path=args.kernel_path)
        sim.step(nsteps)
        got = sim.download()
        sim.close()
    else:
        from nairn_mpm_fea_b200.slab import SlabSim, gather_by_id, partition_particles, slab_bounds
        lo, hi = slab_bounds(prob.depth, 8, 8 + ncell, world)[rank]
        part = partition_particles(prob.particles, prob.horiz, prob.vert, lo, hi)
        ss = SlabSim(prob, part, lo, hi, rank, world, device=local, capacity_factor=1.5)
        ss.step(nsteps)
        got = gather_by_id(ss.download(), n)
        moved = ss.migrated_out
        ss.close()
    if rank != 0 or ref_npz is None or not os.path.exists(ref_npz):
        return None
    z = np.load(ref_npz)
    errs = {}
    for k in PARITY_FIELDS:
        a, b = np.asarray(got[k]), z["b/" + k]
        scale = max(float(np.max(np.abs(z["b/ep"] if k == "wrot" else b))), 1e-300)
        errs[k] = float(np.max(np.abs(a - b))) / scale
    same_elem = bool(np.array_equal(got["in_elem"], z["b/inElem"]))
    return {"max_rel": max(errs.values()), "per_field": errs, "element_ids_identical": same_elem, "steps": nsteps, "particles": int(n),
            "what": "this arm (%s) against the reference arm's CPU sample: %d^3-cell block, same start, %d steps; max |diff| / max |ref| per field"
                    % ("1 GPU" if world == 1 else "%d z-slabs over NCCL" % world, ncell, nsteps)}


def cpu_baseline(ncell, steps, dump=None):
    d = time_reference(ncell, steps, dump=dump)
    if d is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    if "error" in d:
        return {"value": None, "unit": UNIT, "cores": d.get("cores", 0), "kind": "reference", "sample": d["error"]}
    return {"value": d["n"] * d["steps"] / d["run_s"], "unit": UNIT, "cores": d["cores"], "kind": "reference",
            "sample": "reference NairnMPM (oracle/_ref, -O3 -fopenmp) on the same block input (same jittered start and velocity field) at %d^3 cells = %d particles, "
                      "%d steps after 1 warm-up, %d OpenMP threads, %.1f s" % (ncell, d["n"], d["steps"], d["cores"], d["run_s"])}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = time_reference(args.cpu_ncell, args.steps, warm=args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if d is None or "error" in d:
        _JSON_LINES.append(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built or failed: %s" % (d or {}).get("error", "")}))
        return
    v = d["n"] * d["steps"] / d["run_s"]
    sample = ("reference NairnMPM (oracle/_ref) on the bench block input (same jittered start and velocity field) at %d^3 cells = %d particles (bounded sample of the "
              "%s workload), %d steps, %d OpenMP threads" % (args.cpu_ncell, d["n"], args.workload, d["steps"], d["cores"]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * d["run_s"] / d["steps"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (CPU sample: %d^3 cells)" % (args.workload, args.cpu_ncell)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": d["cores"], "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _JSON_LINES.append(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="block8m", choices=["block8m", "block1m", "taylor16m", "taylor2m", "neo8m", "neo1m"])
    ap.add_argument("--ncell", type=int, default=0, help="override block edge in cells (testing)")
    ap.add_argument("--kernel-path", type=int, default=0)
    ap.add_argument("--xpic-order", type=int, default=0, help="block workloads: XPIC(k) particle update (with --fmpm: FMPM(k)); 0 = FLIP")
    ap.add_argument("--fmpm", action="store_true")
    ap.add_argument("--sort-interval", type=int, default=0, help="steps between physical particle sorts (0 = library default)")
    ap.add_argument("--cpu-ncell", type=int, default=50, help="block edge of the CPU sample (50 -> 1M particles)")
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the CPU sample (about 10 s of CPU work on 16 threads at 1M particles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short device-resident runs of BASELINE configs 1-4 (default workload, 1 GPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries the ONE JSON line and nothing else: whatever libraries write to file descriptor 1 meanwhile (NCCL prints its
    # version banner there when NCCL_DEBUG asks for it) goes to stderr; the line itself is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _JSON_LINES:
        sys.stdout.write("\n".join(_JSON_LINES) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
