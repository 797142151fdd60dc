#!/usr/bin/env python
"""bench.py — particle-pushes/s per full PIC step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one lap of projects/pic-turbulence/pic.py:187-221 (push_half_b, B halo,
push, pack+migrate, sort every 5th lap, deposit, J exchange + halo, 3 binomial
filter passes, push_half_b, push_e + add_current, E halo) over a synthetic
uniform thermal pair plasma (BASELINE.json configs[4], "projects/scaling"):
2 species x 16 ppc, theta = 0.3, uniform Bz, tiles of 64^3 cells, a cube of
--cells^3 cells per GPU (default 512^3 = 4.29e9 particles = 137 GB per GPU, the size the
metric is quoted on), GPUs arranged 1 / 2x1x1 / 2x2x1 / 2x2x2 (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Native libraries (NCCL prints its version banner with
# printf) write to file descriptor 1, so the real stdout is kept aside for the result and fd 1 is
# pointed at stderr for the rest of the run.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_RESULT_OUT = None


def emit_result(obj):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def isolate_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)

BYTES_PER_PARTICLE_PUSH = 56.0      # R pos,vel,id (32) + W pos,vel (24)
BYTES_PER_PARTICLE_DEPOSIT = 32.0   # R pos,vel,id
BYTES_PER_PARTICLE_STEP = 88.0      # push + deposit
BYTES_PER_PARTICLE_SORT = 30.0      # amortised over 5 laps
BYTES_PER_CELL_STEP = 300.0


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class Conf:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def gpu_blocks(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def make_conf(args, n_gpus):
    tpg = args.cells // args.tile            # tiles per GPU per axis
    gb = gpu_blocks(n_gpus)
    ppc = args.ppc
    cfl = 0.45
    oppc = 2 * ppc
    q0 = -(cfl ** 2) / (0.5 * oppc * 2.0)    # projects/pic-turbulence/pic.py:49-50 with gamma=c_omp=1, m0=m1=1
    return Conf(n_tiles=[tpg * gb[0], tpg * gb[1], tpg * gb[2]], n_cells_per_tile=[args.tile] * 3, cfl=cfl,
                field_propagator="fdtd2", current_filter="binomial2", q0=q0, m0=1.0, q1=abs(q0), m1=1.0,
                particle_pusher="boris", field_interpolator="linear_1st", current_depositer="zigzag_1st_atomic"), tpg, gb


def binit(conf, ppc, delgam=0.3, sigma=10.0):
    oppc = 2 * ppc
    m0 = conf.m0 * abs(conf.q0)
    gammath = 1.0 + 1.5 * delgam
    return float(np.sqrt(gammath * oppc * m0 * conf.cfl ** 2 * sigma))   # pic.py:64-67


def juttner_synge(rng, n, theta):
    """Isotropic Juettner-Synge momenta by Sobol's rejection method — what runko/sample_thermal_distributions.py:58-127 does
    for theta > 0.2 and what the CUDA arm's device generator (k_inject_thermal) restates: (3, n) float64."""
    u = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.random((4, todo.size))
        uu = -theta * np.log(x[0] * x[1] * x[2])
        eta = -theta * np.log(x[0] * x[1] * x[2] * x[3])
        ok = eta * eta - uu * uu > 1.0
        u[todo[ok]] = uu[ok]
        todo = todo[~ok]
    mu = 2.0 * rng.random(n) - 1.0
    phi = 2.0 * np.pi * rng.random(n)
    st = np.sqrt(np.maximum(0.0, 1.0 - mu * mu))
    return np.stack([u * st * np.cos(phi), u * st * np.sin(phi), u * mu])


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(float(os.environ.get('B2P_CLOCK_PERIOD', '0.1')))

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- reference / CPU arm
class CpuArm:
    """The reference's CPU implementation of the lap on the host cores, one tile per worker at a time
    (the reference's rank-per-core model: serial inside a tile, `OMP_NUM_THREADS=1`).

    kind "reference": the reference's OWN kernel sources (emf::YeeLattice, pic::ParticleContainer)
    compiled from /root/reference into oracle/_ref/libref_kernels.so (-O3 -mavx2 -fopenmp-simd), driven
    tile by tile in corgi's order.  kind "port": the oracle restatement, when that build is absent."""

    def __init__(self, args, n_threads):
        from oracle import reference_build as rbuild
        edge = args.tile                                       # the CUDA arm's tile size (64^3)
        nt = max(1, n_threads)
        tz = 1
        while tz * tz * tz < nt:
            tz += 1
        self.tiles = (tz, tz, max(1, -(-nt // (tz * tz))))
        self.edge, self.n_threads, self.ppc = edge, nt, args.ppc
        conf, _, _ = make_conf(argparse.Namespace(cells=edge, tile=edge, ppc=args.ppc), 1)
        conf.n_tiles = list(self.tiles)
        use_ref = rbuild.available() and rbuild.cpu_ok()
        self.kind = "reference" if use_ref else "port"
        rng = np.random.default_rng(42)
        ncell = edge ** 3
        B = np.zeros((3, edge + 6, edge + 6, edge + 6), np.float32)
        B[2] = binit(conf, args.ppc)
        self.n_part = 0
        if use_ref:
            from concurrent.futures import ThreadPoolExecutor
            self.g = rbuild.RefGrid(conf)
            pool = ThreadPoolExecutor(nt)
            tiles = list(self.g.tiles.values())
            self.g.phase = lambda name: list(pool.map(lambda t: t.op(name), tiles))    # ctypes releases the GIL
            items = [(idx, t) for idx, t in self.g.tiles.items()]
        else:
            from oracle.oracle import OracleGrid
            self.g = OracleGrid(conf)
            T = self.tiles
            items = [((t % T[0], (t // T[0]) % T[1], t // (T[0] * T[1])), t) for t in range(self.g.num_tiles)]
        # one tile's worth of plasma (cell corner + U[0,1)^3, species 1 on top of species 0: pic.py:141-156), repeated in every tile
        ii, jj, kk = np.meshgrid(np.arange(edge), np.arange(edge), np.arange(edge), indexing="ij")
        corner = np.stack([ii.ravel(), jj.ravel(), kk.ravel()]).astype(np.float64)
        pos0 = np.concatenate([corner + rng.random((3, ncell)) for _ in range(args.ppc)], axis=1)
        n = pos0.shape[1]
        vels = [juttner_synge(rng, n, 0.3) for _ in range(2)]      # the CUDA arm's momentum distribution (theta = 0.3)
        for (i, j, k), t in items:
            pos = pos0 + np.array([i * edge, j * edge, k * edge], np.float64)[:, None]
            for sp in range(2):
                vel = vels[sp]
                if use_ref:
                    ids = (np.uint64(sp + 1) << np.uint64(40)) + np.arange(n, dtype=np.uint64)
                    t.set_particles(sp, *pos.astype(np.float32), *vel.astype(np.float32), ids)
                else:
                    self.g.inject(t, sp, *pos, *vel)
                self.n_part += n
            if use_ref:
                t.set_fields(B=B)
            else:
                self.g.set_fields(t, B=B, with_halo=True)
        self.lap = 0
        self.step()                                            # warm-up lap (includes the lap-0 sort)

    def step(self):
        if self.kind == "reference":
            self.g.step_pic(self.lap)
        else:
            self.g.step_pic(self.lap, threads=self.n_threads)
        self.lap += 1

    def sample(self, laps):
        t = self.tiles
        return (f"{t[0]}x{t[1]}x{t[2]} tiles of {self.edge}^3 cells (periodic), 2 species x {self.ppc} ppc, Juettner-Synge theta = 0.3, "
                f"uniform Bz = {self.n_part} particles per lap, {laps} laps of the same lap function, one tile per worker thread")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    arm = CpuArm(args, n_threads)
    for _ in range(args.warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = (time.perf_counter() - t0) / args.steps
    value = arm.n_part / dt
    out = {
        "impl": "reference", "metric": "particle-pushes/s per full PIC step", "value": value, "unit": "particle-pushes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        # the CPU arm times a BOUNDED SAMPLE of that workload (same tile size, ppc, momentum distribution, lap function);
        # `value` is its particle-pushes/s on this sample, i.e. a per-particle rate, not a 512^3 run
        "sample": arm.sample(args.steps),
        "timed_laps": [1 + args.warmup, args.warmup + args.steps],         # lap 0 ran when the arm was set up
        "sort_laps_timed": sum(1 for q in range(1 + args.warmup, 1 + args.warmup + args.steps) if q % 5 == 0),
        "cpu_baseline": {"value": value, "unit": "particle-pushes/s", "cores": n_threads, "kind": arm.kind,
                         "sample": arm.sample(args.steps)},
        "e2e": {"value": value, "unit": "particle-pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(out)


# --------------------------------------------------------------------------- config 3: vacuum EM wave (field-solver roofline)
def owner_map(conf, tpg, gb):
    T = conf.n_tiles
    owner = np.zeros(T[0] * T[1] * T[2], np.int32)
    for k in range(T[2]):
        for j in range(T[1]):
            for i in range(T[0]):
                owner[i + T[0] * (j + T[1] * k)] = (i // tpg) + gb[0] * ((j // tpg) + gb[1] * (k // tpg))
    return owner


def comm_init(L, dist, grid, rank, world, owner):
    from runko_b200._lib import check
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        check(L.b2p_nccl_unique_id(uid.ctypes.data_as(C.c_void_p)))
    lst = [uid.tobytes()]
    dist.broadcast_object_list(lst, src=0)
    uid = np.frombuffer(lst[0], np.uint8).copy()
    check(L.b2p_grid_comm_init(grid._h, rank, world, uid.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p)))


def measure_emf(rb, L, dist, rank, world, local_rank, cells, tile, steps, warmup, with_cpu, with_e2e=True):
    """BASELINE configs[2] ("projects/emf-wave"): FDTD + binomial filter only.  One step = one lap of
    projects/emf-wave/emf.py:48-56 (E halo, push_half_b x2, B halo, push_e) over cells^3 cells per GPU;
    `value` = cell-updates/s.  The three binomial filter passes of a PIC lap are timed separately over the
    same lattices (`per_kernel.filter`)."""
    from runko_b200._lib import check
    tpg, gb = cells // tile, gpu_blocks(world)
    conf = Conf(n_tiles=[tpg * gb[0], tpg * gb[1], tpg * gb[2]], n_cells_per_tile=[tile] * 3, cfl=1.0, field_propagator="fdtd2",
                current_filter="binomial2")
    grid = rb.Grid(conf)
    bi, bj, bk = rank % gb[0], (rank // gb[0]) % gb[1], rank // (gb[0] * gb[1])
    tiles = []
    H = tile + 6
    for k in range(tpg):
        # plane wave k = 2 pi (0,0,1)/10 (emf.py:27-35): Ex(z) at z = k, By(z) at z = k + 1/2 — one slab per tile layer
        z = (bk * tpg + k) * tile + np.arange(H, dtype=np.float64) - 3.0
        E = np.zeros((3, H, H, H), np.float32)
        B = np.zeros((3, H, H, H), np.float32)
        E[0] = np.sin(2 * np.pi * z / 10.0).astype(np.float32)[None, None, :]
        B[1] = np.sin(2 * np.pi * (z + 0.5) / 10.0).astype(np.float32)[None, None, :]
        for i in range(tpg):
            for j in range(tpg):
                t = rb.Tile((bi * tpg + i, bj * tpg + j, bk * tpg + k), conf)
                t.set_fields_f32(E, B, None, with_halo=True)
                grid.add_tile(t)
                tiles.append(t)
    if world > 1:
        comm_init(L, dist, grid, rank, world, owner_map(conf, tpg, gb))
    n_cells_local = cells ** 3

    def barrier():
        rb.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(warmup):
        grid.step_emf()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.b2p_launch_count()
    check(L.b2p_timer_start())
    for _ in range(steps):
        grid.step_emf()
    ms = C.c_float()
    check(L.b2p_timer_stop(C.byref(ms)))
    barrier()
    launches = L.b2p_launch_count() - launches0
    clocks = sampler.stop()
    # per-kernel leg (+ three filter passes per lap, as in a PIC lap)
    check(L.b2p_profile_enable(1))
    prof_steps = 3
    for _ in range(prof_steps):
        grid.step_emf()
        for _ in range(3):
            grid.phase("filter_current")
    rb.sync()
    nk = L.b2p_profile_num_classes()
    pms, pl, pu = np.zeros(nk), np.zeros(nk, np.uint64), np.zeros(nk)
    check(L.b2p_profile_report(pms.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p), pu.ctypes.data_as(C.c_void_p)))
    check(L.b2p_profile_enable(0))
    names = [L.b2p_profile_class_name(k).decode() for k in range(nk)]
    dev_s = ms.value / 1e3
    if dist is not None:
        import torch
        tt = torch.tensor([dev_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s = float(tt[0])
    per_step = dev_s / steps
    peak, peak_src = read_peaks()
    bpu = {"push_b": 36.0, "push_e": 36.0, "filter": 24.0}    # 36 B/cell per sweep, 24 B/cell per filter pass
    per_kernel = {}
    for nm in ("push_b", "push_e", "filter", "halo_fill"):
        k = names.index(nm)
        if pl[k]:
            avg = pms[k] / int(pl[k])
            per_lap = pms[k] / prof_steps
            per_kernel[nm] = {"avg_launch_ms": avg, "launches_per_lap": int(pl[k]) / prof_steps, "ms_per_lap": per_lap}
            if nm in bpu:
                # bytes the launches actually have to move: the two half pushes of B are ONE fused sweep here (36 B/cell,
                # not 2 x 36), so the fused launch is rated on 36 B/cell; the unfused-equivalent figure is kept beside it
                passes = int(pl[k]) / prof_steps
                per_kernel[nm]["GBs"] = bpu[nm] * n_cells_local * passes / (per_lap * 1e-3) / 1e9
                per_kernel[nm]["frac_of_peak"] = per_kernel[nm]["GBs"] / peak
                ref_passes = {"push_b": 2, "push_e": 1, "filter": 3}[nm]
                if ref_passes != passes:
                    per_kernel[nm]["reference_sweeps_replaced_per_lap"] = ref_passes
                    per_kernel[nm]["unfused_equivalent_GBs"] = bpu[nm] * n_cells_local * ref_passes / (per_lap * 1e-3) / 1e9
    top = "push_b"
    achieved = per_kernel[top]["GBs"]
    sweeps = per_kernel["push_b"]["launches_per_lap"] + per_kernel["push_e"]["launches_per_lap"]
    step_bytes = 36.0 * sweeps * n_cells_local                  # the sweeps the shipped lap launches (2 when the half pushes are fused)
    e2e = None
    if with_e2e:
        # e2e: one tile's E,B host round trip + the lap through the per-tile API
        M = rb.comm_mode
        h0, d0 = C.c_uint64(), C.c_uint64()
        rb.sync()
        L.b2p_copy_bytes(C.byref(h0), C.byref(d0))
        t0 = time.perf_counter()
        e2e_steps = max(2, min(steps, 5))
        for s_ in range(e2e_steps):
            t = tiles[s_ % len(tiles)]
            Eh, Bh, _ = t.get_fields_f32(with_halo=True)
            t.set_fields_f32(Eh, Bh, None, with_halo=True)
            if world > 1:
                grid.external_communication(M.emf_E)
            grid.local_communication(M.emf_E)
            for t in tiles: t.push_half_b()
            for t in tiles: t.push_half_b()
            if world > 1:
                grid.external_communication(M.emf_B)
            grid.local_communication(M.emf_B)
            for t in tiles: t.push_e()
        barrier()
        dt = time.perf_counter() - t0
        h1, d1 = C.c_uint64(), C.c_uint64()
        L.b2p_copy_bytes(C.byref(h1), C.byref(d1))
        e2e = {"value": n_cells_local * world / (dt / e2e_steps), "unit": "cell-updates/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int((h1.value - h0.value) / e2e_steps),
               "d2h_bytes_per_step": int((d1.value - d0.value) / e2e_steps),
               "how": "per-tile API; one tile's E and B make a host round trip every step"}
    cpu = None
    if rank == 0 and world == 1 and with_cpu:
        from oracle.oracle import OracleGrid
        nthreads = os.cpu_count() or 1
        edge = 64
        tz = 1
        while tz ** 3 < nthreads:
            tz += 1
        oc = Conf(n_tiles=[tz, tz, max(1, -(-nthreads // (tz * tz)))], n_cells_per_tile=[edge] * 3, cfl=1.0, field_propagator="fdtd2")
        og = OracleGrid(oc)
        c0, nl = time.perf_counter(), 0
        while nl < 3 or (time.perf_counter() - c0 < 10.0 and nl < 200):
            og.step_emf(threads=nthreads)
            nl += 1
        cdt = (time.perf_counter() - c0) / nl
        ncell = edge ** 3 * og.num_tiles
        cpu = {"value": ncell / cdt, "unit": "cell-updates/s", "cores": nthreads, "kind": "port",
               "sample": f"{oc.n_tiles} tiles of {edge}^3 cells, {nl} laps of emf.py:48-56, one tile per worker"}
    out = {"metric": "cell-updates/s per vacuum field lap (BASELINE configs[2], field-solver roofline)",
           "value": n_cells_local * world / per_step, "unit": "cell-updates/s", "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "projects/emf-wave vacuum plane wave (BASELINE configs[2]): E halo, push_half_b x2, B halo, push_e",
                      "cells_per_gpu": f"{cells}^3", "tile": f"{tile}^3", "gpu_blocks": "x".join(map(str, gb)),
                      "l2": "E+B = 24 B/cell x cells far exceed the 126 MB L2; no explicit flush"},
           "clocks": clocks, "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k_push_b_fdtd2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "bytes_per_unit": 36.0,
                        "units_per_launch": n_cells_local, "per_kernel": per_kernel,
                        "step": {"algorithmic_bytes_per_gpu": step_bytes, "achieved_GBs": step_bytes / per_step / 1e9,
                                 "frac_of_peak": step_bytes / per_step / 1e9 / peak,
                                 "note": "36 B/cell per launched sweep; the reference's unfused lap has three sweeps (108 B/cell)",
                                 "unfused_equivalent_GBs": 108.0 * n_cells_local / per_step / 1e9}}}
    if e2e is not None:
        out["e2e"] = e2e
    if cpu is not None:
        out["cpu_baseline"] = cpu
    del tiles, grid
    return out


def run_emf(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    import runko_b200 as rb
    from runko_b200._lib import check
    L = rb.lib()
    check(L.b2p_init(local_rank))
    cells = args.cells if args.cells != 512 else 1024
    tile = args.tile if args.tile != 64 else 128
    out = measure_emf(rb, L, dist, rank, world, local_rank, cells, tile, args.steps, args.warmup, not args.no_cpu_baseline)
    if rank == 0:
        emit_result(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



# --------------------------------------------------------------------------- configs[1] beam and configs[3] shock
STENCIL_X = dict(stencil_x_delta=-0.13512912586798165, stencil_x_gamma=0.016100341594971104, stencil_x_beta_p1=0.022350417588794424,
                 stencil_x_beta_p2=0.022350417588794424, stencil_x_zeta_p1=0.012206124727660462, stencil_x_zeta_p2=0.012206124727660462)


def named_workload(name, n_gpus, cells_arg=None):
    """Shapes and physics of BASELINE configs[1] / configs[3] (per GPU; more GPUs extend the box along gpu_blocks)."""
    gb = gpu_blocks(n_gpus)
    if name == "beam":
        # projects/pic-beam-instabilities/beam.py:46-137: one species, two cold counter-streaming beams (gamma_b = 3) of 8 ppc each,
        # thin-z 3-D box, filter off as shipped (enable_filter = False)
        cells, tile = (cells_arg or 1024, cells_arg or 1024, 6), (64, 64, 6)
        cfl, skin, n0 = 0.45, 10.0, 16
        gmean = 3.0
        q = -((cfl / skin) ** 2) * gmean / n0
        conf = Conf(cfl=cfl, field_propagator="fdtd2", q0=q, m0=1.0, particle_pusher="boris", field_interpolator="linear_1st",
                    current_depositer="zigzag_1st_atomic", current_filter=None)
        desc = {"workload": "projects/pic-beam-instabilities two-stream / filamentation (BASELINE configs[1]): 1 species, 2 cold beams "
                            "gamma_b = 3 along +-x, 2 x 8 ppc, no current filter (beam.py:60)", "ppc_total": 16}
        species, ppc_total = 1, 16
    else:
        # projects/pic-shock/pic.py + 3d_shock_mini.ini physics: sigma = 3, upstream gamma = 5, theta = 1e-5, faraday + stencil +
        # binomial2_unrolled + linear_1st_unrolled, 4 filter passes, wall at x = 15, moving injector
        cells, tile = (cells_arg or 2048, (cells_arg or 2048) // 8, (cells_arg or 2048) // 8), (64, 64, 64)
        cfl, skin, ppc = 0.45, 10.0, 4
        oppc = 2 * ppc
        q = -((cfl / skin) ** 2) * 5.0 / (0.5 * oppc * (1.0 + 1.0))   # pic.py:47-50 (omp = cfl/c_omp, m0 = m1 = 1)
        conf = Conf(cfl=cfl, field_propagator="stencil", q0=q, m0=1.0, q1=abs(q), m1=1.0, particle_pusher="faraday",
                    field_interpolator="linear_1st_unrolled", current_depositer="zigzag_1st_atomic", current_filter="binomial2_unrolled",
                    prealloc_per_species=int(2.0 * ppc * tile[0] * tile[1] * tile[2]), **STENCIL_X)
        desc = {"workload": "projects/pic-shock reflecting-wall shock with moving injector (BASELINE configs[3]): 2 species x 4 ppc upstream "
                            "(8 ppc), sigma = 3, gamma_up = 5, faraday + stencil + binomial2_unrolled x4 + linear_1st_unrolled, wall at x = 15",
                "ppc_total": 8}
        species, ppc_total = 2, 8
    tpg = tuple(cells[d] // tile[d] for d in range(3))
    conf.n_tiles = [tpg[d] * gb[d] for d in range(3)]
    conf.n_cells_per_tile = list(tile)
    desc.update({"cells_per_gpu": "x".join(map(str, cells)), "tile": "x".join(map(str, tile)), "gpu_blocks": "x".join(map(str, gb)),
                 "fp": "fp32, IEEE div/sqrt, NO FMA contraction on either arm",
                 "l2": "particle streams (32 B x particles) far exceed the 126 MB L2; no explicit flush"})
    return conf, tpg, gb, desc, species, ppc_total


class ShockState:
    """what projects/pic-shock/pic.py keeps on the host around the lap: upstream fields, wall, edge BCs, injector front"""

    def __init__(self, conf, Lx):
        self.Lx, self.walloc, self.gamma_up, self.theta, self.ppc = float(Lx), 15.0, 5.0, 1e-5, 4
        self.beta = float(np.sqrt(1.0 - 1.0 / self.gamma_up ** 2))
        oppc, sigma = 2 * self.ppc, 3.0
        m0 = abs(conf.q0)
        self.binit = float(np.sqrt(self.gamma_up * oppc * 0.5 * conf.cfl ** 2 * m0 * 2.0 * sigma))      # pic.py:61
        self.E_up, self.B_up = (0.0, -self.beta * self.binit, 0.0), (0.0, 0.0, self.binit)          # b_proj = (0, 0, 1)
        self.injloc = self.walloc + 0.6 * (self.Lx - self.walloc)       # a shock that has been running: 60 % of the box is filled
        self.n_inj = 5

    def inject_front(self, grid, lap, cfl, seed):
        if lap == 0 or lap % self.n_inj:
            return
        stride = self.n_inj * cfl                                        # runko/moving_injector.py
        left, right = max(self.injloc - self.beta * stride, self.walloc), self.injloc + 1.0 * stride
        if right >= self.Lx - 10.0:
            return
        for sp in range(2):
            grid.inject_drifting_stripe(sp, self.ppc, self.theta, self.gamma_up, -1, left, right, seed=seed + lap)
        self.injloc = right


def run_named(args):
    """--workload beam | shock: the same JSON contract as the headline workload (value = alive particles pushed per second over whole
    laps, roofline of the dominant kernel class, e2e through the per-tile API with a host round trip, cpu_baseline = the oracle
    port on a bounded sample of the same physics)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    import runko_b200 as rb
    from runko_b200._lib import check
    L = rb.lib()
    check(L.b2p_init(local_rank))
    name = args.workload
    conf, tpg, gb, desc, n_species, ppc_total = named_workload(name, world, args.cells if args.cells != 512 else None)
    tile = conf.n_cells_per_tile
    grid = rb.Grid(conf)
    bi, bj, bk = rank % gb[0], (rank // gb[0]) % gb[1], rank // (gb[0] * gb[1])
    Lx = conf.n_tiles[0] * tile[0]
    shock = ShockState(conf, Lx) if name == "shock" else None
    tiles = []
    if shock:
        shape = (3, tile[0] + 6, tile[1] + 6, tile[2] + 6)
        E0, B0 = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
        for c in range(3):
            E0[c], B0[c] = shock.E_up[c], shock.B_up[c]
        wall = rb.reflector_wall(walloc=shock.walloc)
        cbc = rb.edge_bc(direction=0, side=0, position=shock.walloc, E_components=0b110, B_components=0, J_components=0b111)
        ubc = rb.edge_bc(direction=0, side=1, position=Lx - 5.0, Ex=shock.E_up[0], Ey=shock.E_up[1], Ez=shock.E_up[2],
                         Bx=shock.B_up[0], By=shock.B_up[1], Bz=shock.B_up[2], J_components=0b111)
    for i in range(tpg[0]):
        for j in range(tpg[1]):
            for k in range(tpg[2]):
                t = rb.PicTile((bi * tpg[0] + i, bj * tpg[1] + j, bk * tpg[2] + k), conf)
                if shock:
                    t.set_fields_f32(E0, B0, None, with_halo=True)
                    t.register_reflector_wall(wall); t.register_edge_bc(cbc); t.register_edge_bc(ubc)
                grid.add_tile(t)
                tiles.append(t)
    if world > 1:
        # block decomposition with non-cubic tile counts per GPU
        T = conf.n_tiles
        owner = np.zeros(T[0] * T[1] * T[2], np.int32)
        for k in range(T[2]):
            for j in range(T[1]):
                for i in range(T[0]):
                    owner[i + T[0] * (j + T[1] * k)] = (i // tpg[0]) + gb[0] * ((j // tpg[1]) + gb[1] * (k // tpg[2]))
        comm_init(L, dist, grid, rank, world, owner)
        grid._multi = True
    if name == "beam":
        for sign in (+1, -1):
            grid.inject_drifting_stripe(0, 8, 1e-5, 3.0, sign, 0.0, float(Lx), seed=42)      # both beams on the same positions (beam.py:127)
    else:
        for sp in range(2):
            grid.inject_drifting_stripe(sp, shock.ppc, shock.theta, shock.gamma_up, -1, shock.walloc, shock.injloc, seed=42)
    rb.sync()

    def barrier():
        rb.sync()
        if dist is not None:
            dist.barrier()

    def step(lap):
        if shock:
            grid.step_shock(lap, n_filter_passes=4)
            shock.inject_front(grid, lap, conf.cfl, 1000)
        else:
            grid.step_pic(lap)

    for m in (rb.comm_mode.emf_E, rb.comm_mode.emf_B):
        if world > 1:
            grid.external_communication(m)
        grid.local_communication(m)
    lap = 0
    for _ in range(args.warmup):
        step(lap); lap += 1
    barrier()
    alive0 = grid.alive_counts()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.b2p_launch_count()
    barrier()
    check(L.b2p_timer_start())
    for _ in range(args.steps):
        step(lap); lap += 1
    ms = C.c_float()
    check(L.b2p_timer_stop(C.byref(ms)))
    barrier()
    launches = L.b2p_launch_count() - launches0
    clocks = sampler.stop()
    alive1 = grid.alive_counts()
    n_part_local = 0.5 * (float(np.sum(alive0)) + float(np.sum(alive1)))     # alive particles pushed per lap (the shock's count grows)
    # per-kernel leg
    prof_steps = 5
    check(L.b2p_set_option(b"push_streams", 1)); check(L.b2p_set_option(b"sort_streams", 0))
    check(L.b2p_profile_enable(1))
    for _ in range(prof_steps):
        step(lap); lap += 1
    rb.sync()
    nk = L.b2p_profile_num_classes()
    pms, pl, pu = np.zeros(nk), np.zeros(nk, np.uint64), np.zeros(nk)
    check(L.b2p_profile_report(pms.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p), pu.ctypes.data_as(C.c_void_p)))
    check(L.b2p_profile_enable(0))
    check(L.b2p_set_option(b"push_streams", 1)); check(L.b2p_set_option(b"sort_streams", 2))
    names = [L.b2p_profile_class_name(k).decode() for k in range(nk)]
    dev_s = ms.value / 1e3
    tot_part = n_part_local
    if dist is not None:
        import torch
        tt = torch.tensor([dev_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s = float(tt[0])
        tp = torch.tensor([n_part_local], dtype=torch.float64)
        dist.all_reduce(tp)
        tot_part = float(tp[0])
    per_step = dev_s / args.steps
    peak, peak_src = read_peaks()
    top = int(np.argmax(pms))
    share = {names[k]: round(float(pms[k] / max(pms.sum(), 1e-9)), 4) for k in np.argsort(-pms)[:8] if pms[k] > 0}
    fused = pl[names.index("deposit")] == 0 or pms[names.index("deposit")] < 0.2 * pms[names.index("push")]
    push_ms = pms[names.index("push")] / prof_steps
    alive_now = float(np.sum(grid.alive_counts()))
    achieved = BYTES_PER_PARTICLE_STEP * alive_now / (push_ms * 1e-3) / 1e9 if fused else BYTES_PER_PARTICLE_PUSH * alive_now / (push_ms * 1e-3) / 1e9
    n_cells_local = int(np.prod(tpg)) * int(np.prod(tile))
    step_bytes = (BYTES_PER_PARTICLE_STEP + BYTES_PER_PARTICLE_SORT) * n_part_local + BYTES_PER_CELL_STEP * n_cells_local
    roofline = {"bound": "hbm", "kernel": "push+deposit (fused k_push)" if fused else "push (k_push; deposit separate where the wall reflects)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "bytes_per_unit": BYTES_PER_PARTICLE_STEP if fused else BYTES_PER_PARTICLE_PUSH, "units": "alive particles per lap",
                "ms_per_lap": push_ms, "share_of_step": share, "dominant_class": names[top],
                "step": {"algorithmic_bytes_per_gpu": step_bytes, "achieved_GBs": step_bytes / per_step / 1e9,
                         "frac_of_peak": step_bytes / per_step / 1e9 / peak, "frac_of_8TBs": step_bytes / per_step / 8e12}}
    if args.profile and rank == 0:
        for k in np.argsort(-pms):
            if pl[k]:
                print(f"  {names[k]:16s} {pms[k] / prof_steps:9.3f} ms/step  {int(pl[k]) // prof_steps:6d} launches/step", file=sys.stderr)
    checks = {"alive_before": [int(v) for v in alive0], "alive_after": [int(v) for v in alive1]}
    if name == "beam":
        checks["particle_number_conserved"] = bool(np.array_equal(alive0, alive1))
    # ---- e2e: the lap tile by tile through the reference-facing API, one tile's particles making a host round trip per step
    e2e = None
    if not args.no_e2e:
        M = rb.comm_mode
        multi = world > 1

        def comm(m, local=None):
            if multi:
                grid.external_communication(m)
            grid.local_communication(m if local is None else local)

        def each(method, *a):
            for t in tiles:
                getattr(t, method)(*a)

        def lap_api(lp):
            each("push_half_b")
            if shock: each("apply_edge_bcs", M.emf_B)
            comm(M.emf_B)
            each("push_particles")
            if shock: each("reflect_particles")
            each("pack_outgoing_particles"); comm(M.pic_particle)
            if lp % 5 == 0: each("sort_particles")
            each("deposit_current"); comm(M.emf_J, M.emf_J_exchange); comm(M.emf_J)
            if shock:
                each("apply_edge_bcs", M.emf_J)
                for q in range(4):
                    if q > 0 and q % 3 == 0: comm(M.emf_J)
                    each("filter_current")
                each("apply_edge_bcs", M.emf_J)
            each("push_half_b")
            if shock: each("apply_edge_bcs", M.emf_B)
            comm(M.emf_B)
            each("push_e")
            if shock: each("apply_edge_bcs", M.emf_E)
            each("add_current")
            if shock: each("apply_edge_bcs", M.emf_E)
            comm(M.emf_E)
            if shock: each("advance_reflector_walls")
            return grid.energies()

        steps = max(2, min(args.steps, 5))
        h0, d0 = C.c_uint64(), C.c_uint64()
        barrier()
        L.b2p_copy_bytes(C.byref(h0), C.byref(d0))
        t0 = time.perf_counter()
        for s_ in range(steps):
            t = tiles[(len(tiles) // 2 + s_) % len(tiles)]
            for sp in range(n_species):
                st = t.get_particles(sp, alive_only=False)
                t.set_particles_raw(sp, *st)
            lap_api(lap + s_)
        barrier()
        dt = time.perf_counter() - t0
        lap += steps
        h1, d1 = C.c_uint64(), C.c_uint64()
        L.b2p_copy_bytes(C.byref(h1), C.byref(d1))
        if dist is not None:
            import torch
            tt = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        e2e = {"value": tot_part / (dt / steps), "unit": "particle-pushes/s", "steps": steps,
               "h2d_bytes_per_step": int((h1.value - h0.value) / steps), "d2h_bytes_per_step": int((d1.value - d0.value) / steps),
               "how": "lap driven tile-by-tile through the PicTile API; every step one tile's particle state makes a host round trip "
                      "(get_particles -> set_particles) and the energy diagnostics are read back; wall clock, max over ranks"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = named_cpu_baseline(name, os.cpu_count() or 1)
    if rank == 0:
        out = {"metric": "particle-pushes/s per full PIC step", "value": tot_part / per_step, "unit": "particle-pushes/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": desc,
               "particles_per_gpu": n_part_local, "cell_updates_per_s": n_cells_local * world / per_step,
               "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "checks": checks}
        if e2e is not None:
            out["e2e"] = e2e
        if cpu is not None:
            out["cpu_baseline"] = cpu
        emit_result(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def named_cpu_baseline(name, n_threads):
    """The oracle port (kind "port") on a bounded sample of the same physics: one tile per worker thread, the workload's own tile
    shape and particle density, the same lap function (for the shock: wall tile row + upstream rows, edge BCs, 4 filter passes)."""
    from oracle.oracle import OracleGrid
    conf, tpg, gb, desc, n_species, ppc_total = named_workload(name, 1)
    tile = conf.n_cells_per_tile
    rng = np.random.default_rng(42)
    if name == "beam":
        ty = max(1, int(np.sqrt(n_threads)))
        conf.n_tiles = [max(1, n_threads // ty), ty, 1]
    else:
        tile = [32, 32, 32]                                  # 64^3 tiles x 8 ppc x all cores would run for minutes per lap
        conf.n_cells_per_tile = tile
        conf.prealloc_per_species = int(2.0 * 4 * 32 ** 3)
        conf.n_tiles = [max(2, n_threads // 2), 2, 1]
    T = conf.n_tiles
    og = OracleGrid(conf)
    Lx = T[0] * tile[0]
    shock = ShockState(conf, Lx) if name == "shock" else None
    ii, jj, kk = np.meshgrid(np.arange(tile[0]), np.arange(tile[1]), np.arange(tile[2]), indexing="ij")
    corner = np.stack([ii.ravel(), jj.ravel(), kk.ravel()]).astype(np.float64)
    n_part = 0

    def drifting(n, theta, Gamma, sign):
        u = np.sqrt(theta) * rng.standard_normal((3, n))
        g = np.sqrt(1.0 + np.sum(u * u, axis=0))
        beta = np.sqrt(1.0 - 1.0 / Gamma ** 2)
        flip = -beta * u[0] / g > rng.random(n)
        u[0, flip] = -u[0, flip]
        u[0] = sign * Gamma * (u[0] + beta * g)
        return u

    for t in range(og.num_tiles):
        i, j, k = t % T[0], (t // T[0]) % T[1], t // (T[0] * T[1])
        org = np.array([i * tile[0], j * tile[1], k * tile[2]], np.float64)[:, None]
        if name == "beam":
            pos = np.concatenate([corner + rng.random(corner.shape) for _ in range(8)], axis=1) + org
            for sign in (+1, -1):
                og.inject(t, 0, *pos, *drifting(pos.shape[1], 1e-5, 3.0, sign))
                n_part += pos.shape[1]
        else:
            shape = (3, tile[0] + 6, tile[1] + 6, tile[2] + 6)
            E0, B0 = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
            for c in range(3):
                E0[c], B0[c] = shock.E_up[c], shock.B_up[c]
            og.set_fields(t, E0, B0, None, with_halo=True)
            import runko_b200.tiles as _t
            og.register_reflector_wall(t, _t.reflector_wall(walloc=shock.walloc))
            og.register_edge_bc(t, _t.edge_bc(direction=0, side=0, position=shock.walloc, E_components=0b110, B_components=0, J_components=0b111))
            og.register_edge_bc(t, _t.edge_bc(direction=0, side=1, position=Lx - 5.0, Ex=shock.E_up[0], Ey=shock.E_up[1], Ez=shock.E_up[2],
                                              Bx=shock.B_up[0], By=shock.B_up[1], Bz=shock.B_up[2], J_components=0b111))
            pos = np.concatenate([corner + rng.random(corner.shape) for _ in range(shock.ppc)], axis=1) + org
            keep = (pos[0] >= shock.walloc) & (pos[0] < shock.injloc)
            pos = pos[:, keep]
            for sp in range(2):
                og.inject(t, sp, *pos, *drifting(pos.shape[1], shock.theta, shock.gamma_up, -1))
                n_part += pos.shape[1]

    def lap_fn(lap):
        if name == "beam":
            og.step_pic(lap, threads=n_threads)
            return
        ph, lc = (lambda nm: og.phase(nm, threads=n_threads)), og.local_communication
        ph("push_half_b"); ph("apply_edge_bcs_B"); lc(2)
        ph("push_particles"); ph("reflect_particles"); ph("pack_outgoing_particles"); lc(3)
        if lap % 5 == 0:
            ph("sort_particles")
        ph("deposit_current"); lc(6); lc(0); ph("apply_edge_bcs_J")
        for q in range(4):
            if q > 0 and q % 3 == 0:
                lc(0)
            ph("filter_current")
        ph("apply_edge_bcs_J")
        ph("push_half_b"); ph("apply_edge_bcs_B"); lc(2)
        ph("push_e"); ph("apply_edge_bcs_E"); ph("add_current"); ph("apply_edge_bcs_E"); lc(1)
        ph("advance_reflector_walls")

    for m in (1, 2):
        og.local_communication(m)
    lap_fn(0)
    c0, nl = time.perf_counter(), 0
    while nl < 3 or (time.perf_counter() - c0 < 12.0 and nl < 50):
        lap_fn(1 + nl)
        nl += 1
    cdt = (time.perf_counter() - c0) / nl
    return {"value": n_part / cdt, "unit": "particle-pushes/s", "cores": n_threads, "kind": "port",
            "sample": f"{T[0]}x{T[1]}x{T[2]} tiles of {tile[0]}x{tile[1]}x{tile[2]} cells, {n_part} particles, {nl} laps of the workload's lap function, one tile per worker thread"}


def workload_config(args, n_gpus):
    gb = gpu_blocks(n_gpus)
    return {"workload": "projects/scaling uniform thermal pair plasma (BASELINE configs[4] physics), weak scaling",
            "cells_per_gpu": f"{args.cells}^3", "tile": f"{args.tile}^3", "species": 2, "ppc_per_species": args.ppc,
            "particles_per_gpu": 2 * args.ppc * args.cells ** 3, "gpu_blocks": "x".join(map(str, gb)),
            "lap": "pic-turbulence/pic.py:187-221, sort every 5th lap, fdtd2 + boris + linear_1st + zigzag_1st_atomic + 3x binomial2",
            "fp": "fp32, IEEE div/sqrt, NO FMA contraction on either arm (nvcc -fmad=false; CPU arm -ffp-contract=off): the convention "
                  "the bit-exact parity tests pin (the reference's own presets would let the compiler contract)",
            "l2": "per-step working set (32 B x particles >= 17 GB) far exceeds the 126 MB L2; no explicit flush"}


def multi_gpu_equals_single(rb, L, dist, rank, world, gb):
    """Parity evidence a multi-GPU run can carry: a small periodic grid (same physics, 2^3 tiles of 16^3 cells per rank,
    2 x 4 ppc) is advanced (a) by every rank alone, all tiles local, and (b) by the N ranks together, each owning its
    block, halos / currents / particles crossing NCCL.  After lap 0 (push, migration, sort, deposit, field update) every
    owned container must hold bit-identical particles in identical slots and B must agree bit for bit; after 6 laps the
    particle ids per (tile, species) still agree and the fields agree within the deposit tolerance (atomic order)."""
    tpb, edge, ppc = 2, 16, 4
    conf, _, _ = make_conf(argparse.Namespace(cells=tpb * edge, tile=edge, ppc=ppc), world)
    T = conf.n_tiles
    own = lambda i, j, k: (i // tpb) + gb[0] * ((j // tpb) + gb[1] * (k // tpb))      # noqa: E731
    grids, tiles = [], []
    for multi in (False, True):
        g = rb.Grid(conf)
        ts = {}
        for i in range(T[0]):
            for j in range(T[1]):
                for k in range(T[2]):
                    if not multi or own(i, j, k) == rank:
                        t = rb.PicTile((i, j, k), conf)
                        g.add_tile(t)
                        ts[(i, j, k)] = t
        if multi:
            comm_init(L, dist, g, rank, world, owner_map(conf, tpb, gb))
        g.set_uniform_B(0.0, 0.0, binit(conf, ppc))
        g.inject_thermal(ppc, 0.3, seed=7)
        for m in (rb.comm_mode.emf_E, rb.comm_mode.emf_B):
            if multi:
                g.external_communication(m)
            g.local_communication(m)
        grids.append(g)
        tiles.append(ts)

    def compare(exact):
        ok_p, ok_b, ferr = True, True, 0.0
        for idx, tm in tiles[1].items():
            ts = tiles[0][idx]
            for sp in range(2):
                a, b = ts.get_particles(sp, alive_only=False), tm.get_particles(sp, alive_only=False)
                if exact:
                    alive = a[6] != np.uint64(0xFFFFFFFFFFFFFFFF)
                    ok_p = ok_p and len(a[6]) == len(b[6]) and bool(np.array_equal(a[6], b[6])) and \
                        all(np.array_equal(a[c][alive].view(np.uint32), b[c][alive].view(np.uint32)) for c in range(6))
                else:
                    ok_p = ok_p and bool(np.array_equal(np.sort(ts.get_ids(sp)), np.sort(tm.get_ids(sp))))
            fa, fb = ts.get_fields_f32(with_halo=False), tm.get_fields_f32(with_halo=False)
            ok_b = ok_b and bool(np.array_equal(fa[1].view(np.uint32), fb[1].view(np.uint32)))
            for x, y in zip(fa, fb):
                ferr = max(ferr, float(np.max(np.abs(x - y)) / max(float(np.max(np.abs(x))), 1e-30)))
        return ok_p, ok_b, ferr

    for g in grids:
        g.step_pic(0)
    p0, b0, _ = compare(True)
    for lap in range(1, 6):
        for g in grids:
            g.step_pic(lap)
    p5, _, f5 = compare(False)
    import torch
    flags = torch.tensor([int(p0), int(b0), int(p5)], dtype=torch.int64)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    err = torch.tensor([f5], dtype=torch.float64)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    del tiles, grids
    return {"grid": f"{T[0]}x{T[1]}x{T[2]} tiles of {edge}^3 cells, 2 x {ppc} ppc, {world} ranks vs 1 rank",
            "lap0_particles_bit_exact": bool(flags[0]), "lap0_B_bit_exact": bool(flags[1]),
            "lap5_ids_per_tile_equal": bool(flags[2]), "lap5_max_field_rel_err": float(err[0]),
            "ok": bool(flags[0]) and bool(flags[1]) and bool(flags[2]) and float(err[0]) <= 1e-3}


# --------------------------------------------------------------------------- CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=512, help="cube edge of cells per GPU")
    ap.add_argument("--tile", type=int, default=64)
    ap.add_argument("--ppc", type=int, default=16, help="particles per cell per species")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-emf", action="store_true", help="skip the emf-wave sub-record of the default workload")
    ap.add_argument("--profile", action="store_true", help="per-kernel-class timing table on stderr")
    ap.add_argument("--workload", default="scaling", choices=["scaling", "emf-wave", "beam", "shock"],
                    help="scaling: BASELINE configs[4], the headline (default); emf-wave: configs[2], fields only, cell-updates/s; "
                         "beam: configs[1] (1024x1024x6, 16 ppc); shock: configs[3] (2048x256x256, 8 ppc, wall + injector)")
    args = ap.parse_args()
    isolate_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "emf-wave":
        return run_emf(args)
    if args.workload in ("beam", "shock"):
        return run_named(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    n_gpus = world

    import runko_b200 as rb
    from runko_b200._lib import check
    L = rb.lib()
    check(L.b2p_init(local_rank))

    conf, tpg, gb = make_conf(args, n_gpus)
    grid = rb.Grid(conf)
    # this rank's block of tiles
    bi, bj, bk = rank % gb[0], (rank // gb[0]) % gb[1], rank // (gb[0] * gb[1])
    tiles = []
    for i in range(tpg):
        for j in range(tpg):
            for k in range(tpg):
                t = rb.PicTile((bi * tpg + i, bj * tpg + j, bk * tpg + k), conf)
                grid.add_tile(t)
                tiles.append(t)
    if world > 1:
        comm_init(L, dist, grid, rank, world, owner_map(conf, tpg, gb))
    grid.set_uniform_B(0.0, 0.0, binit(conf, args.ppc))
    grid.inject_thermal(args.ppc, 0.3, seed=42)
    n_part_local = 2 * args.ppc * args.cells ** 3
    n_cells_local = args.cells ** 3
    rb.sync()

    def barrier():
        rb.sync()
        if dist is not None:
            dist.barrier()

    # prelude: sync E,B halos (pic.py:177-185)
    for m in (rb.comm_mode.emf_E, rb.comm_mode.emf_B):
        if world > 1:
            grid.external_communication(m)
        grid.local_communication(m)
    lap = 0
    alive0 = en0 = None
    for w in range(args.warmup):
        if w == args.warmup - 1:
            # the reference state of the invariant checks is taken one lap before the timed region: the push of the lap
            # that follows an energy read keeps the library's kinetic-energy account (+8 % of that push), and `value`
            # is the bare loop of laps
            alive0, en0 = grid.alive_counts(), grid.energies()
        grid.step_pic(lap)
        lap += 1
    barrier()
    if en0 is None:
        alive0, en0 = grid.alive_counts(), grid.energies()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.b2p_launch_count()
    barrier()
    t0 = time.perf_counter()
    check(L.b2p_timer_start())
    step_wall, step_blocked = [], []
    hw0, hw1 = C.c_double(), C.c_double()
    for _ in range(args.steps):
        L.b2p_host_wait_ms(C.byref(hw0))
        ts = time.perf_counter()
        grid.step_pic(lap)
        step_wall.append(time.perf_counter() - ts)
        L.b2p_host_wait_ms(C.byref(hw1))
        step_blocked.append((hw1.value - hw0.value) * 1e-3)
        lap += 1
    ms = C.c_float()
    check(L.b2p_timer_stop(C.byref(ms)))
    barrier()
    wall = time.perf_counter() - t0
    launches = L.b2p_launch_count() - launches0
    clocks = sampler.stop()

    # ---- per-kernel leg: the same laps once more (one sort cycle), with the worker streams off so
    # that every kernel class runs alone on the library stream and a CUDA-event pair around each
    # launch measures that launch only (in the timed region above up to 4 tiles' kernels overlap)
    prof_steps = 5
    check(L.b2p_set_option(b"push_streams", 1))
    check(L.b2p_set_option(b"sort_streams", 0))
    check(L.b2p_profile_enable(1))
    for _ in range(prof_steps):
        grid.step_pic(lap)
        lap += 1
    rb.sync()
    nk = L.b2p_profile_num_classes()
    pms, pl, pu = np.zeros(nk), np.zeros(nk, np.uint64), np.zeros(nk)
    check(L.b2p_profile_report(pms.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p), pu.ctypes.data_as(C.c_void_p)))
    check(L.b2p_profile_enable(0))
    check(L.b2p_set_option(b"push_streams", 1))      # the defaults of common.cuh: Tuning
    check(L.b2p_set_option(b"sort_streams", 2))
    names = [L.b2p_profile_class_name(k).decode() for k in range(nk)]
    barrier()

    dev_s = ms.value / 1e3
    t_max = dev_s
    if dist is not None:
        import torch
        tt = torch.tensor([dev_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt[0])
    per_step = t_max / args.steps
    value = n_part_local * world / per_step

    # ---- roofline of the dominant kernel class (largest share of the timed region) ----
    peak, peak_src = read_peaks()
    fused = pl[names.index("deposit")] == 0      # push + deposit fused: credited with both phases' algorithmic bytes
    bytes_per_unit = {"push": BYTES_PER_PARTICLE_STEP if fused else BYTES_PER_PARTICLE_PUSH, "deposit": BYTES_PER_PARTICLE_DEPOSIT, "sort_gather": 64.0,
                      "detect_leavers": 20.0, "sort_count": 28.0, "sort_place": 64.0, "filter": 24.0, "push_b": 36.0,
                      "push_e": 48.0, "zero": 4.0, "halo_fill": 0.0, "J_exchange": 0.0, "nodal_means": 56.0}
    top = int(np.argmax(pms))
    share = {names[k]: round(float(pms[k] / max(pms.sum(), 1e-9)), 4) for k in np.argsort(-pms)[:8] if pms[k] > 0}
    avg_ms = pms[top] / max(int(pl[top]), 1)
    units_per_launch = pu[top] / max(int(pl[top]), 1)            # container slots (dead ones included)
    if names[top] in ("push", "deposit"):
        # credited per ALIVE particle: the containers carry dead slots (leavers, slack) that the kernel skips
        units_per_launch = float(np.sum(grid.alive_counts())) * prof_steps / max(int(pl[top]), 1)
    achieved = bytes_per_unit.get(names[top], 0.0) * units_per_launch / (avg_ms * 1e-3) / 1e9
    step_bytes = (BYTES_PER_PARTICLE_STEP + BYTES_PER_PARTICLE_SORT) * n_part_local + BYTES_PER_CELL_STEP * n_cells_local
    kname = "push+deposit (fused k_push)" if (fused and names[top] == "push") else names[top]
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (not measured live)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_push_traffic.json")
    if names[top] == "push" and fused and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["traffic_bytes_per_launch"] * units_per_launch / tj["alive_particles_per_launch"]
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture scaled to this launch's alive particles, profiles/r02_push_traffic.json)",
                "dram_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "units": "alive particles per launch",
                "peak_source": peak_src,
                "how": f"CUDA events around every launch of the class over {prof_steps} further laps (one sort cycle) run on the "
                       "library stream alone (worker streams off); `step` below is the timed region itself",
                "avg_launch_ms": avg_ms, "launches": int(pl[top]), "bytes_per_unit": bytes_per_unit.get(names[top], 0.0),
                "units_per_launch": units_per_launch, "share_of_step": share,
                "step": {"algorithmic_bytes_per_gpu": step_bytes, "achieved_GBs": step_bytes / per_step / 1e9,
                         "frac_of_peak": step_bytes / per_step / 1e9 / peak, "frac_of_8TBs": step_bytes / per_step / 8e12}}
    if args.profile and rank == 0:
        print("  host wall time per step_pic call [ms]:", " ".join(f"{1e3 * v:.1f}" for v in step_wall), file=sys.stderr)
        print("  ... of which enqueueing (not blocked in a stream synchronisation) [ms]:",
              " ".join(f"{1e3 * (v - b):.1f}" for v, b in zip(step_wall, step_blocked)), file=sys.stderr)
        for k in np.argsort(-pms):
            if pl[k]:
                print(f"  {names[k]:16s} {pms[k] / prof_steps:9.3f} ms/step  {int(pl[k]) // prof_steps:6d} launches/step", file=sys.stderr)

    # ---- size-independent invariants at the full bench size (no oracle at 4.3e9 particles): particle number,
    # energy budget, and the sort contract on one container (sorted by cell key, dead slots last, idempotent)
    alive1, en1 = grid.alive_counts(), grid.energies()
    tot = [alive0, alive1]
    if dist is not None:
        import torch
        tt = torch.tensor(np.stack(tot).astype(np.int64))
        dist.all_reduce(tt)
        tot = [tt[0].numpy(), tt[1].numpy()]
    m0 = abs(conf.q0)
    e_tot = [e[0] + e[1] + m0 * float(np.sum(e[2])) for e in (en0, en1)]
    t0_ = tiles[0]
    t0_.sort_particles()
    keys = t0_.sort_keys(0).astype(np.int64)
    ids_a = t0_.get_particles(0, alive_only=False)[6].copy()
    t0_.sort_particles()
    ids_b = t0_.get_particles(0, alive_only=False)[6]
    dead = ids_a == np.uint64(0xFFFFFFFFFFFFFFFF)
    n_alive0 = int(np.count_nonzero(~dead))
    checks = {"particles_before": [int(v) for v in tot[0]], "particles_after": [int(v) for v in tot[1]],
              "particle_number_conserved": bool(np.array_equal(tot[0], tot[1])),
              "energy_drift_over_timed_and_profiled_laps": e_tot[1] / e_tot[0] - 1.0,
              "container_sorted_by_cell_key": bool(np.all(np.diff(keys) >= 0)),
              "dead_slots_last": bool(not np.any(dead[:n_alive0]) and np.all(dead[n_alive0:])),
              "sort_idempotent": bool(np.array_equal(ids_a, ids_b))}

    # ---- e2e: whole job through the reference-facing per-tile API with host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, rb, L, conf, grid, tiles, world, dist)

    if world > 1:
        checks["multi_gpu_equals_single"] = multi_gpu_equals_single(rb, L, dist, rank, world, gb)

    # ---- field-solver sub-record (BASELINE configs[2]) on the same box, after the PIC state is freed
    emf = None
    if not args.no_emf:
        del t0_, tiles, grid
        rb.sync()
        emf_cells = min(args.cells, 512)
        emf = measure_emf(rb, L, dist, rank, world, local_rank, emf_cells, min(128, emf_cells), 10, 3, with_cpu=False, with_e2e=False)
        emf = {k: emf[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "gpu_launches", "roofline")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_threads = os.cpu_count() or 1
        arm = CpuArm(args, n_threads)
        c0 = time.perf_counter()
        nl = 0
        while nl < 3 or (time.perf_counter() - c0 < 12.0 and nl < 50):
            arm.step()
            nl += 1
        cdt = (time.perf_counter() - c0) / nl
        cpu = {"value": arm.n_part / cdt, "unit": "particle-pushes/s", "cores": n_threads, "kind": arm.kind, "sample": arm.sample(nl)}

    if rank == 0:
        out = {"metric": "particle-pushes/s per full PIC step", "value": value, "unit": "particle-pushes/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(args, world),
               "timed_laps": [args.warmup, args.warmup + args.steps - 1],      # laps are numbered from 0; multiples of 5 sort
               "sort_laps_timed": sum(1 for q in range(args.warmup, args.warmup + args.steps) if q % 5 == 0),
               "cell_updates_per_s": n_cells_local * world / per_step, "wall_ms_per_step": wall / args.steps * 1e3,
               "host_enqueue_ms_per_step": (sum(step_wall) - sum(step_blocked)) / args.steps * 1e3,
               "kernel_ms_per_step_single_stream": float(pms.sum() / prof_steps),
               "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "checks": checks}
        if e2e is not None:
            out["e2e"] = e2e
        if emf is not None:
            out["emf_wave"] = emf
        if cpu is not None:
            out["cpu_baseline"] = cpu
        emit_result(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, rb, L, conf, grid, tiles, world, dist):
    """Whole-job end to end through the reference-facing tile API (runko_b200.PicTile methods, the
    calls runko/simulation.py makes per tile) starting and ending in HOST memory: every timed step
    (i) re-uploads one tile's particle state from pinned-equivalent host arrays through the public
    setter, (ii) runs the lap tile by tile, (iii) reads the per-lap energy diagnostics and one tile's
    particle state back to the host.  Bytes are counted by the library (b2p_copy_bytes)."""
    from runko_b200._lib import check
    M = rb.comm_mode
    multi = world > 1

    def ext(m):
        if multi:
            grid.external_communication(m)

    def lap_via_tile_api(lap):
        for t in tiles: t.push_half_b()
        ext(M.emf_B); grid.local_communication(M.emf_B)
        for t in tiles: t.push_particles()
        for t in tiles: t.pack_outgoing_particles()
        ext(M.pic_particle); grid.local_communication(M.pic_particle)
        if lap % 5 == 0:
            for t in tiles: t.sort_particles()
        for t in tiles: t.deposit_current()
        ext(M.emf_J); grid.local_communication(M.emf_J_exchange)
        ext(M.emf_J); grid.local_communication(M.emf_J)
        for t in tiles: t.filter_current()
        ext(M.emf_J); grid.local_communication(M.emf_J)
        for t in tiles: t.filter_current()
        for t in tiles: t.filter_current()
        for t in tiles: t.push_half_b()
        ext(M.emf_B); grid.local_communication(M.emf_B)
        for t in tiles: t.push_e()
        for t in tiles: t.add_current()
        ext(M.emf_E); grid.local_communication(M.emf_E)
        return grid.energies()                                   # io_average_* (D2H)

    steps = max(2, min(args.steps, 5))
    # page-locked host buffers for the particle round trip (registered once, outside the timed region)
    cap = max(t.container_size(sp) for t in tiles[:steps] for sp in range(2)) + (1 << 20)
    host = [[np.empty(cap, np.float32) for _ in range(6)] + [np.empty(cap, np.uint64)] for _ in range(2)]
    for sp in range(2):
        for a in host[sp]:
            check(L.b2p_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes))
    h0, d0 = C.c_uint64(), C.c_uint64()
    rb.sync()
    if dist is not None:
        dist.barrier()
    L.b2p_copy_bytes(C.byref(h0), C.byref(d0))
    t0 = time.perf_counter()
    lap = 1000   # keeps lap % 5 phase: laps 1000..: sort on the first
    for s in range(steps):
        t = tiles[s % len(tiles)]
        state = [t.get_particles(sp, alive_only=False, out=host[sp]) for sp in range(2)]   # D2H of one tile
        for sp in range(2):
            t.set_particles_raw(sp, *state[sp])                                   # H2D of one tile
        lap_via_tile_api(lap + s)
    rb.sync()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    h1, d1 = C.c_uint64(), C.c_uint64()
    L.b2p_copy_bytes(C.byref(h1), C.byref(d1))
    if args.profile and world == 1:
        # where an end-to-end step spends its time (one further step, a synchronisation after every section; not timed above)
        t = tiles[steps % len(tiles)]
        ta = time.perf_counter()
        state = [t.get_particles(sp, alive_only=False, out=host[sp]) for sp in range(2)]
        rb.sync(); tb = time.perf_counter()
        for sp in range(2):
            t.set_particles_raw(sp, *state[sp])
        rb.sync(); tc = time.perf_counter()
        lap_via_tile_api(lap + steps)
        rb.sync(); td = time.perf_counter()
        grid.energies()
        te = time.perf_counter()
        print(f"  e2e step sections [ms]: tile state D2H {1e3 * (tb - ta):.1f}, H2D {1e3 * (tc - tb):.1f}, "
              f"lap through the tile API + diagnostics {1e3 * (td - tc):.1f}, diagnostics alone {1e3 * (te - td):.1f}", file=sys.stderr)
    for sp in range(2):
        for a in host[sp]:
            L.b2p_host_unregister(a.ctypes.data_as(C.c_void_p))
    if dist is not None:
        import torch
        tt = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
    n_part = 2 * args.ppc * args.cells ** 3 * world
    return {"value": n_part / (dt / steps), "unit": "particle-pushes/s", "steps": steps,
            "h2d_bytes_per_step": int((h1.value - h0.value) / steps), "d2h_bytes_per_step": int((d1.value - d0.value) / steps),
            "how": "lap driven tile-by-tile through the PicTile API (as runko/simulation.py does); every step one tile's whole "
                   "particle state makes a host round trip through page-locked buffers (get_particles -> set_particles) and the "
                   "energy diagnostics are read back; wall clock, max over ranks"}


if __name__ == "__main__":
    main()
