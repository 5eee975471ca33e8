#!/usr/bin/env python3
"""bench.py -- element matrices/s of the batched stiffness integration (BASELINE.json metric) on N B200s.

Headline workload (BASELINE.json configs[1]): 3D Hex8 Block mesh 200^3 cells with its interior points jittered by 0.2 h,
integ::mat_10_bdb with isotropic elasticity (E=480, nu=1/3), 2x2x2 Gauss, f64.  A "step" is one pass of the hot path over
one GPU's 8,000,000 cells with mesh and output resident in HBM (36.9 GB of element matrices per pass: far larger than L2,
no cache flush needed).  At N>1 every rank owns one 200^3 slab (a contiguous cell range of an N-slab-tall Block mesh);
there is no collective on the data path (the logical result is the concatenation of the ranks' outputs): "weak" scaling.

Keys beyond the base contract:
  parity        GPU vs the CPU oracle on >= 10^4 seeded sample cells of the jittered full-size mesh, inside this run.
  roofline      the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json); algorithmic bytes =
                4,664 B/cell; fp64 sub-object: 37,264 flop/cell against the DFMA peak measured in-process.
  strong        the SAME 200^3 mesh split by partition() over the N ranks (north_star: one mesh, contiguous ranges).
  configs       (N = 1) one record per BASELINE.json config -- Qua4 1000^2, Hex20 100^3 27-point (bulged), Tet10 5e6 mass +
                source, mixed Tri3/Qua8 1e7 axisymmetric -- each with elements/s, the kernels that ran, roofline fractions and
                oracle parity on >= 10^4 sampled cells.
  cpu_baseline  the C oracle (a port of the reference's per-cell Rust loop; the reference needs cargo and cannot be built
                here) timed on this host's cores over a bounded sample of the same workload (rank 0, N=1 only).
  e2e           the same metric through the reference-facing call with HOST buffers, >= 1e6 cells per GPU per step: the mesh
                goes host->device from pinned memory, the kernels run, the element matrices come back into pinned host
                memory.  K is symmetric here (D is major-symmetric), so the primary e2e line ships the packed upper triangles
                (gl_args.packed: 300 instead of 576 doubles per element); `e2e_full` ships the full 24x24 blocks.

`--impl reference` times the reference's CPU path (the oracle port, all host threads) on a bounded sample of the same
workload and prints the same JSON line with "impl": "reference"; it imports nothing of the product package.
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

NDIV = (200, 200, 200)
YOUNG, POISSON = 480.0, 1.0 / 3.0
BYTES_PER_CELL = 4664.0   # 4608 out + 32 connectivity + 24.4 coordinates
FLOPS_PER_CELL = 37264.0  # structured two-stage contraction, no symmetry halving
METRIC = "element matrices/sec (Hex8 K=int B.D.B, f64)"
UNIT = "elements/s"
PARITY_SAMPLE = 10_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ndiv", type=int, nargs=3, default=list(NDIV), help="cells per direction of one rank's slab")
    ap.add_argument("--e2e-layers", type=int, default=25, help="z-layers of the slab used for the host-buffer (e2e) leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="cells in the CPU baseline sample (0 = auto, ~10 s)")
    ap.add_argument("--config-scale", type=float, default=1.0, help="cell-count scale of the per-config records")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    nx, ny, nz = args.ndiv
    return {
        "workload": f"hex8 Block mesh {nx}x{ny}x{nz} cells per GPU ({nx * ny * nz} cells), interior points jittered by 0.2 h, "
                    f"mat_10_bdb isotropic elasticity E=480 nu=1/3, 2x2x2 Gauss (BASELINE.json configs[1])",
        "cells_per_gpu": nx * ny * nz,
        "global_cells": nx * ny * nz * n_gpus,
        "partition": f"{n_gpus} contiguous cell ranges (one {nx}x{ny}x{nz} slab per rank), no collectives",
        "l2": "inputs+outputs (36.9 GB/step) far larger than the 126 MB L2; no flush needed",
    }


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram bytes per cell of a kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), or None.  This is
    a constant read from the committed profile, not a live counter."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(key)
        except Exception:
            return None
    return None


def bind_to_gpu_numa_node(device_index):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates next are
    placed (first touch) in memory local to that GPU's PCIe root.  Best effort: returns the node or None."""
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                           capture_output=True, text=True, timeout=20)
        bus = r.stdout.strip().splitlines()[0].strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:  # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        if node < 0:
            return None
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation (oracle port) -- nothing of the product package is imported here
# ---------------------------------------------------------------------------------------------------------------------------

def hex8_lattice(nx, ny, nz, lengths=(1.0, 1.0, 1.0)):
    """A Hex8 lattice mesh as plain numpy arrays (cell id = i + nx (j + ny k), local nodes in the order of
    src/shapes/hex8.rs:122-131, lattice-order point ids): the reference arm's input."""
    ax, ay = nx + 1, ny + 1
    e = np.arange(nx * ny * nz, dtype=np.int64)
    i, j, k = e % nx, (e // nx) % ny, e // (nx * ny)
    loc = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.stack([(i + a) + ax * ((j + b) + ay * (k + c)) for a, b, c in loc], axis=1).astype(np.uint64)
    q = np.arange(ax * ay * (nz + 1), dtype=np.int64)
    coords = np.stack([(q % ax) / nx * lengths[0], ((q // ax) % ay) / ny * lengths[1], (q // (ax * ay)) / nz * lengths[2]], axis=1)
    kinds = np.full(e.size, 17, dtype=np.uint8)  # GeoKind::Hex8 in VALUES order
    conn_off = (np.arange(e.size + 1, dtype=np.uint64) * 8)
    return coords, kinds, conn_off, conn.reshape(-1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po

    nx, ny, nz = args.ndiv
    layers = max(1, min(nz, 8))
    coords, kinds, conn_off, conn = hex8_lattice(nx, ny, layers, (1.0, 1.0, layers / nz))
    rng = np.random.default_rng(12345)  # same distortion as the GPU arm: interior points jittered by 0.2 h
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    interior = np.all((coords > lo + 1e-12) & (coords < hi - 1e-12), axis=1)
    coords[interior] += rng.uniform(-0.2 / nx, 0.2 / nx, size=(int(interior.sum()), 3))
    dd = np.asarray(po.lin_elasticity(YOUNG, POISSON, False, False), dtype=np.float64).ravel(order="F")
    ncell = kinds.size
    threads = po.ncores()
    sample = min(args.cpu_sample or ncell, ncell)
    out = np.empty(sample * 576)

    def leg(n):
        t0 = time.perf_counter()
        po.mesh_integ("mat_10_bdb", 3, coords, kinds, conn_off, conn, 0, n, coef=dd, nthreads=threads, out=out[:n * 576])
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        leg(max(sample // 8, 1000))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        leg(sample)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} cells per step (first {layers} z-layers of the jittered slab), OpenMP over cells, "
                                   f"{threads} threads; the Rust original cannot be built here (no cargo), this is its C restatement"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------------

def time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_config_record(torch, ctx, po, idx, scale, hbm_peak, fp64_peak, threads):
    """One BASELINE.json config on this GPU: device-resident timing (CUDA events), kernels that ran, roofline fractions,
    oracle parity on a seeded sample of the distorted mesh."""
    from gemlab_b200 import Coef, CommonArgs, integ
    from tools import workloads as wl

    t_build = time.perf_counter()
    cfg = wl.build_config(idx, scale)
    mesh = cfg["mesh"]
    dm = ctx.upload(mesh)
    t_build = time.perf_counter() - t_build
    a = CommonArgs(axisymmetric=cfg["axisymmetric"])
    outs = [torch.empty(integ.out_total(dm, op, None, a), dtype=torch.float64, device="cuda") for op, _ in cfg["cases"]]
    kernels = []

    def step(record=False):
        # every case of the configuration from ONE gl_integ_fused call: the library shares the geometry pass where it has a kernel
        # for it (mass + source; conductivity + flux; on Qua8 cells stiffness + conductivity + flux) and runs the rest in turn
        items = [(op, Coef.uniform(c), o) for (op, c), o in zip(cfg["cases"], outs)]
        if len(items) == 1:
            integ.mat_10_bdb(ctx, dm, None, a, items[0][1], out=items[0][2])
        else:
            integ.fused(ctx, dm, None, a, items)
        if record:
            kernels.extend(ctx.last_kernel().split(" + "))

    launches0 = ctx.launch_count()
    step(record=True)
    launches = ctx.launch_count() - launches0
    ms = time_steps(torch, step, steps=5, warmup=2)
    ctx.sync()
    worst, ncmp = 0.0, 0
    for (op, c), o in zip(cfg["cases"], outs):
        kw = {"axisymmetric": True} if cfg["axisymmetric"] else {}
        err, ncmp = wl.parity_sample(po, mesh, op, o.cpu().numpy(), c, PARITY_SAMPLE, nthreads=threads, **kw)
        worst = max(worst, err)
    els = mesh.ncell / (ms * 1e-3)
    rec = {
        "config": cfg["name"], "cells": int(mesh.ncell), "ms_per_pass": ms, "elements_per_s": els, "kernels": kernels,
        "gpu_launches_per_pass": int(launches),
        "roofline": {"bound": "hbm", "achieved": cfg["bytes"] * els / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": cfg["bytes"] * els / 1e9 / hbm_peak, "kernel_ms": ms, "algorithmic_bytes_per_cell": cfg["bytes"]},
        "fp64": {"achieved": cfg["flops"] * els / 1e12, "peak_dfma_measured": fp64_peak, "unit": "TFLOP/s",
                 "frac": cfg["flops"] * els / 1e12 / fp64_peak, "algorithmic_flops_per_cell": cfg["flops"]},
        "parity": {"max_rel_err": worst, "cells_sampled": ncmp, "against": "oracle (C restatement of the reference's per-cell path)",
                   "tolerance": 1e-12},
        "host_mesh_build_and_upload_s": t_build,
    }
    dm.free()
    del outs
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import gemlab_b200 as g
    from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, integ

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; keep stdout for the single JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = Context(local_rank)
    # rank r owns the r-th 200^3 slab of an N-slab-tall block.  The slab is generated on the device (gl_mesh_block: the
    # reference's Block::subdivide cell order and first-touch point numbering); its interior points are then jittered by
    # 0.2 h so that no two element matrices are alike.
    nx, ny, nz = args.ndiv
    dmesh = ctx.block(g.GeoKind.Hex8, (nx, ny, nz), xmin=[0.0, 0.0, float(rank)], xmax=[1.0, 1.0, float(rank) + 1.0])
    host = ctx.download(dmesh)
    host = g.jitter_interior_points(host, 0.2 / nx, 12345 + rank)
    dmesh.update_coords(host.coords)
    ncell = dmesh.ncell
    dd = g.mandel_matrix(g.lin_elasticity(YOUNG, POISSON, False))
    coef = Coef.uniform(dd)
    cargs = CommonArgs()
    out = torch.empty(ncell * 576, dtype=torch.float64, device="cuda")

    def step():
        integ.mat_10_bdb(ctx, dmesh, (0, ncell), cargs, coef, out=out)

    # nvidia-smi takes a few hundred ms to deliver its first sample and then one every 100 ms, while K timed steps may last less than
    # that: the sampler is started BEFORE the warm-up, the warm-up is extended (same kernel, untimed) until a sample has arrived, and
    # if none lands inside the timed region the same kernel keeps running (untimed) until the next one does -- every sample used was
    # taken under the load of the timed kernel, right around or inside the region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()
    if rank == 0 and sampler.proc:
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            step()
            ctx.sync()
    kernel_name = ctx.last_kernel()
    barrier()
    # The timed region: EXACTLY K steps between two events, barrier + synchronize on both sides.  If no clock sample landed inside it
    # (nvidia-smi's real period on these boxes is 150 - 300 ms, K steps last ~130 ms), the measurement is repeated -- same K steps,
    # same barriers, at most four attempts -- and the attempt that contains a sample is the one reported (`timed_attempts`).
    attempts = 0
    while True:
        attempts += 1
        launches0 = ctx.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        n_before = len(sampler.rows)
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        n_after = len(sampler.rows)
        launches = ctx.launch_count() - launches0
        again = 1.0 if (rank == 0 and sampler.proc and n_after == n_before and attempts < 4) else 0.0
        if max_over_ranks(again) == 0.0:
            break
    if rank == 0 and sampler.proc and n_after == n_before:
        t_wait = time.perf_counter()
        while len(sampler.rows) == n_after and time.perf_counter() - t_wait < 1.0:
            step()
            ctx.sync()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if rank == 0:
        sampler.rows = sampler.rows[max(n_before - 1, 0):]
        clocks = sampler.stop()
        clocks["samples_inside_timed_region"] = n_after - n_before
        clocks["timed_attempts"] = attempts
    else:
        clocks = None
    ctx.sync()  # surfaces device-side failures (zero determinant)
    ms_max = max_over_ranks(ms)
    value = world * ncell * args.steps / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps

    # ---- strong scaling: ONE 200^3 mesh, rank r integrates partition(ncell, N, r) (north_star) ---------------------------
    lo, hi = g.partition(ncell, world, rank)
    for _ in range(3):
        integ.mat_10_bdb(ctx, dmesh, (lo, hi), cargs, coef, out=out)
    barrier()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sv0.record()
    for _ in range(args.steps):
        integ.mat_10_bdb(ctx, dmesh, (lo, hi), cargs, coef, out=out)
    sv1.record()
    torch.cuda.synchronize()
    barrier()
    strong_ms = max_over_ranks(sv0.elapsed_time(sv1))
    strong = {"value": ncell * args.steps / (strong_ms * 1e-3), "unit": UNIT, "ms_per_step": strong_ms / args.steps,
              "scaling": "strong", "global_cells": ncell,
              "partition": f"one {nx}x{ny}x{nz} mesh, rank r integrates cells [floor(r n/N), floor((r+1) n/N))"}
    step()  # the full slab again: the parity sample below reads it
    ctx.sync()

    # ---- parity at full size: >= 10^4 seeded sample cells of the jittered mesh against the oracle (rank 0) ---------------
    parity = None
    if rank == 0:
        from oracle import pyoracle as po
        from tools import workloads as wl

        sel = np.sort(np.random.default_rng(1).choice(ncell, size=min(PARITY_SAMPLE, ncell), replace=False))
        got = out.view(ncell, 576)[torch.from_numpy(sel).cuda()].cpu().numpy()
        sub_mesh = host.permuted(sel)
        ref = po.mesh_integ("mat_10_bdb", 3, sub_mesh.coords, sub_mesh.kinds, sub_mesh.conn_off, sub_mesh.conn, coef=dd,
                            nthreads=po.ncores()).reshape(sel.size, 576)
        err = np.max(np.abs(got - ref), axis=1) / np.max(np.abs(ref), axis=1)
        k0 = got[0].reshape(24, 24, order="F")
        parity = {"max_rel_err": float(err.max()), "cells_sampled": int(sel.size), "tolerance": 1e-12,
                  "against": "oracle (C restatement of the reference's per-cell path), jittered full-size mesh",
                  "symmetry_rel": float(np.max(np.abs(k0 - k0.T)) / np.max(np.abs(k0))),
                  "rigid_translation_rel": max(float(np.max(np.abs(k0 @ np.tile(np.eye(3)[i], 8)))) for i in range(3)) / float(np.max(np.abs(k0)))}
        del wl

    # ---- e2e legs: host buffers through the public call, >= 1e6 cells per GPU per step ------------------------------------
    e2e = e2e_full = None
    sub = None
    layers = max(1, min(args.e2e_layers, nz))
    if not args.no_e2e or (world == 1 and not args.no_cpu):
        sub = Block.new_cube(1.0).set_ndiv([nx, ny, layers]).subdivide(GeoKind.Hex8)
        sub.coords[:, 2] = sub.coords[:, 2] * (layers / nz) + float(rank)
        sub = g.jitter_interior_points(sub, 0.2 / nx, 777 + rank)
    if not args.no_e2e:
        # the step's inputs live in PINNED host memory (allocated after the NUMA binding above: local to this GPU's PCIe root)
        pinned = []
        for name in ("coords", "kinds", "markers", "conn_off", "conn"):
            t_pin = torch.from_numpy(getattr(sub, name)).clone().pin_memory()
            pinned.append(t_pin)
            setattr(sub, name, t_pin.numpy())
        n_e2e = sub.ncell
        host_out = torch.empty(n_e2e * 576, dtype=torch.float64).pin_memory()
        h2d = sub.coords.nbytes + sub.conn.nbytes + dd.nbytes  # what gl_mesh_upload copies: coordinates + 64-bit point ids (usize)
        e2e_steps = max(2, min(args.steps, 4))

        def e2e_leg(packed):
            per_cell = 300 if packed else 576
            a = CommonArgs(packed=packed)
            buf = host_out[:n_e2e * per_cell]

            def e2e_step():
                dm = ctx.upload(sub)
                integ.mat_10_bdb(ctx, dm, None, a, coef, out=buf)
                dm.free()

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            return world * n_e2e * e2e_steps / dt, n_e2e * per_cell * 8, buf

        def e2e_parity(buf, packed):
            from oracle import pyoracle as po

            nchk = min(2000, n_e2e)
            ref = po.mesh_integ("mat_10_bdb", 3, sub.coords, sub.kinds, sub.conn_off, sub.conn, 0, nchk, coef=dd,
                                nthreads=po.ncores()).reshape(nchk, 576)
            got = buf[:nchk * (300 if packed else 576)].numpy()
            got = integ.unpack_symmetric(got, 24) if packed else got.reshape(nchk, 576)
            return float(np.max(np.max(np.abs(got - ref), axis=1) / np.max(np.abs(ref), axis=1))), nchk

        what = (f"{nx}x{ny}x{layers} cells per GPU per step ({n_e2e} cells): gl_mesh_upload (pinned host coords + 64-bit connectivity) + "
                f"gl_mat_10_bdb into a pinned host buffer; chunked, kernels hidden behind the device->host copy")
        # what the host link gives: one 1 GiB pinned device->host copy per rank, all ranks at once (the e2e legs move 2.4 / 4.6 KB
        # per element over this link and are bound by it; on a box whose ranks share host bandwidth the sum does not scale with N)
        link_src = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
        link_dst = host_out[: 1 << 27]
        link_dst.copy_(link_src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            link_dst.copy_(link_src, non_blocking=True)
        torch.cuda.synchronize()
        link_gbs_rank = 3 * link_src.numel() * 8 / (time.perf_counter() - t0) / 1e9
        barrier()
        link_all = world * 3 * link_src.numel() * 8 / max_over_ranks(time.perf_counter() - t0) / 1e9
        del link_src
        v, d2h, buf = e2e_leg(packed=True)
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "output": "packed upper triangles (gl_args.packed): 300 doubles per element -- K is symmetric (D is major-symmetric)",
               "workload": what, "numa_node": numa_node,
               "d2h_link_gbs": {"this_rank_alone_in_its_share": link_gbs_rank, "all_ranks_concurrently": link_all,
                                "note": "measured pinned device->host copy rate; value * d2h_bytes / n_cells is bounded by it"},
               "frac_of_link": v * (d2h / n_e2e) / 1e9 / link_all}
        if rank == 0:
            e2e["parity_max_rel_err"], e2e["parity_cells"] = e2e_parity(buf, True)
        v, d2h, buf = e2e_leg(packed=False)
        e2e_full = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "output": "full 24x24 blocks: 576 doubles per element (exactly the reference's output)", "workload": what}
        if rank == 0:
            e2e_full["parity_max_rel_err"], e2e_full["parity_cells"] = e2e_parity(buf, False)
        del host_out, buf

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        kernel_ms = ms / args.steps  # rank 0's own kernel time per launch (one kernel per step)
        achieved = BYTES_PER_CELL * ncell / (kernel_ms * 1e-3) / 1e9
        traffic_pc = ncu_traffic("hex8_mat_10_bdb_dram_bytes_per_cell")
        fp64_peak = ctx.fp64_peak("dfma", 8192)
        dmma_peak = ctx.fp64_peak("dmma", 2048)
        fp64_ach = FLOPS_PER_CELL * ncell / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": None if traffic_pc is None else traffic_pc * ncell,
            "traffic_source": "profiles/ncu_traffic.json (committed ncu --set full capture; a constant, not a live counter)",
            "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_cell": BYTES_PER_CELL,
            "fp64": {"achieved": fp64_ach, "peak_dfma_measured": fp64_peak, "peak_dmma_measured": dmma_peak,
                     "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak, "algorithmic_flops_per_cell": FLOPS_PER_CELL,
                     "note": "Hex8 stiffness sits on the roofline ridge (8.0 flop/B): both fractions are reported"},
        }
        cpu = None
        configs = None
        if world == 1:
            from oracle import pyoracle as po

            threads = po.ncores()
            if not args.no_cpu:
                def cpu_leg(n, nthreads):
                    o = np.empty(n * 576)
                    t0 = time.perf_counter()
                    po.mesh_integ("mat_10_bdb", 3, sub.coords, sub.kinds, sub.conn_off, sub.conn, 0, n, coef=dd, nthreads=nthreads, out=o)
                    return n / (time.perf_counter() - t0), time.perf_counter() - t0

                rate1, _ = cpu_leg(min(20000, sub.ncell), 1)
                sample = args.cpu_sample or int(min(sub.ncell, max(20000, rate1 * threads * 8.0)))
                rate, dt = cpu_leg(sample, threads)
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"first {sample} cells of the jittered e2e sub-slab, {dt:.1f} s, OpenMP over cells with {threads} threads "
                                 f"(1 thread, the reference's own execution model: {rate1:.0f} elements/s)",
                       "single_thread_value": rate1}
            if not args.no_configs:
                del out
                torch.cuda.empty_cache()
                configs = [{"config": workload_config(args, 1)["workload"], "cells": int(ncell), "ms_per_pass": kernel_ms,
                            "elements_per_s": ncell / (kernel_ms * 1e-3), "kernels": [kernel_name],
                            "roofline": {k: roofline[k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms")},
                            "parity": parity}]
                for idx in (0, 2, 3, 4):
                    configs.append(run_config_record(torch, ctx, po, idx, args.config_scale, hbm_peak, fp64_peak, threads))
                configs.sort(key=lambda r: r["config"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "e2e_full": e2e_full, "gpu_launches": int(launches), "kernel": kernel_name,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "strong": strong, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
