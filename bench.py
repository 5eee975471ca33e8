#!/usr/bin/env python3
"""bench.py -- element matrices/s of the batched stiffness integration (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[1]): 3D Hex8 Block mesh 200^3 cells, integ::mat_10_bdb with isotropic elasticity
(E=480, nu=1/3), 2x2x2 Gauss, f64.  A "step" is one pass of the hot path over one GPU's 8,000,000 cells with mesh and
output resident in HBM (36.9 GB of element matrices per pass, i.e. far larger than L2, so no cache flush is needed).
At N>1 every rank owns one 200^3 slab (a contiguous cell range of an N-slab-tall Block mesh); there is no collective on
the data path (the logical result is the concatenation of the ranks' outputs), so scaling is "weak".

Keys beyond the base contract:
  roofline      the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes =
                4,664 B/cell; fp64 sub-object: algorithmic flops (37,264/cell) against the DFMA peak
                measured in-process by gl_bench_fp64_peak.
  cpu_baseline  the C oracle (a port of the reference's per-cell Rust loop; the reference needs cargo and cannot be built
                here) timed on this host's cores over a bounded sample of the same workload (rank 0, N=1 only).
  e2e           the same metric through the reference-facing call with HOST buffers: per step the mesh (coords +
                connectivity) goes host->device, the kernels run and the element matrices come back device->host.

`--impl reference` times the reference's CPU path (the oracle port, all host threads) on a bounded sample of the same
workload and prints the same JSON line with "impl": "reference".
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ndiv", type=int, nargs=3, default=list(NDIV), help="cells per direction of one rank's slab")
    ap.add_argument("--e2e-layers", type=int, default=50, help="z-layers of the slab used for the host-buffer (e2e) leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="cells in the CPU baseline sample (0 = auto, ~10 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    nx, ny, nz = args.ndiv
    return {
        "workload": f"hex8 Block mesh {nx}x{ny}x{nz} cells per GPU ({nx * ny * nz} cells), mat_10_bdb isotropic "
                    f"elasticity E=480 nu=1/3, 2x2x2 Gauss (BASELINE.json configs[1])",
        "cells_per_gpu": nx * ny * nz,
        "global_cells": nx * ny * nz * n_gpus,
        "partition": f"{n_gpus} contiguous cell ranges (one 200^3 slab per rank), no collectives",
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


def ncu_traffic_per_cell():
    """dram bytes per cell of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("hex8_mat_10_bdb_dram_bytes_per_cell")
        except Exception:
            return None
    return None


def slab_mesh(args, rank):
    from gemlab_b200 import Block, GeoKind

    nx, ny, nz = args.ndiv
    mesh = Block.new_cube(1.0).set_ndiv([nx, ny, nz]).subdivide(GeoKind.Hex8)
    mesh.coords[:, 2] += float(rank)  # rank r owns the r-th slab of an N-slab-tall block
    return mesh


def cpu_leg(args, mesh, dd, sample_cells, threads):
    """Times the oracle's mesh loop (set_pad gather + integ::mat_10_bdb per cell) on `threads` host threads."""
    from oracle import pyoracle as po

    n = min(sample_cells, mesh.ncell)
    out = np.empty(n * 576)
    t0 = time.perf_counter()
    po.mesh_integ("mat_10_bdb", 3, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, 0, n, coef=dd, nthreads=threads,
                  out=out)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; cargo/rustc are not in the image
    so the Rust original cannot be built) with all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gemlab_b200 as g
    from oracle import pyoracle as po

    nx, ny, nz = args.ndiv
    layers = max(1, min(nz, 8))
    from gemlab_b200 import Block, GeoKind

    mesh = Block.new_cube(1.0).set_ndiv([nx, ny, layers]).subdivide(GeoKind.Hex8)
    mesh.coords[:, 2] *= layers / nz
    dd = g.mandel_matrix(g.lin_elasticity(YOUNG, POISSON, False))
    threads = po.ncores()
    sample = args.cpu_sample or mesh.ncell
    for _ in range(args.warmup):
        cpu_leg(args, mesh, dd, max(sample // 8, 1000), threads)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, n, _ = cpu_leg(args, mesh, dd, sample, threads)
        done += n
    dt = time.perf_counter() - t0
    value = done / dt
    cfg = workload_config(args, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{min(sample, mesh.ncell)} cells per step (first {layers} z-layers of the slab), "
                                   f"OpenMP over cells, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import gemlab_b200 as g
    from gemlab_b200 import Coef, CommonArgs, Context, integ

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
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

    ctx = Context(local_rank)
    # rank r owns the r-th 200^3 slab of an N-slab-tall block; the slab is generated on the device (gl_mesh_block: the
    # reference's Block::subdivide cell order and first-touch point numbering), so no rank builds an 8M-cell host mesh
    nx, ny, nz = args.ndiv
    dmesh = ctx.block(g.GeoKind.Hex8, (nx, ny, nz), xmin=[0.0, 0.0, float(rank)], xmax=[1.0, 1.0, float(rank) + 1.0])
    ncell = dmesh.ncell
    dd = g.mandel_matrix(g.lin_elasticity(YOUNG, POISSON, False))
    coef = Coef.uniform(dd)
    cargs = CommonArgs()
    out = torch.empty(ncell * 576, dtype=torch.float64, device="cuda")

    def step():
        integ.mat_10_bdb(ctx, dmesh, (0, ncell), cargs, coef, out=out)

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ctx.sync()  # surfaces device-side failures (zero determinant)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * ncell * args.steps / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps

    # size-independent correctness properties at full size (tests/ hold the oracle comparisons at small sizes)
    k0 = out[:576].cpu().numpy().reshape(24, 24, order="F")
    sym = float(np.max(np.abs(k0 - k0.T)) / np.max(np.abs(k0)))
    rigid = max(float(np.max(np.abs(k0 @ np.tile(np.eye(3)[i], 8)))) for i in range(3)) / float(np.max(np.abs(k0)))
    same = float(torch.max(torch.abs(out.view(ncell, 576)[-1] - out.view(ncell, 576)[0])).item() / np.max(np.abs(k0)))
    checks = {"symmetry_rel": sym, "rigid_translation_rel": rigid, "first_vs_last_cell_rel": same}

    # ---- e2e leg: host buffers through the public call -------------------------------------------------------
    # host sub-slab (first `layers` z-layers of this rank's slab): the e2e inputs and the CPU-baseline sample.  The pinned
    # result buffer is 4.6 KB per cell, so the layer count shrinks with the number of ranks sharing the host's memory.
    e2e = None
    sub = None
    layers = max(1, min(max(4, args.e2e_layers // world), nz))
    if not args.no_e2e or (world == 1 and not args.no_cpu):
        from gemlab_b200 import Block, GeoKind

        sub = Block.new_cube(1.0).set_ndiv([nx, ny, layers]).subdivide(GeoKind.Hex8)
        sub.coords[:, 2] = sub.coords[:, 2] * (layers / nz) + float(rank)
    if not args.no_e2e:
        # the step's inputs live in PINNED host memory (the mesh arrays are re-homed into page-locked buffers)
        pinned = []
        for name in ("coords", "kinds", "markers", "conn_off", "conn"):
            t_pin = torch.from_numpy(getattr(sub, name)).clone().pin_memory()
            pinned.append(t_pin)
            setattr(sub, name, t_pin.numpy())
        n_e2e = sub.ncell
        host_out = torch.empty(n_e2e * 576, dtype=torch.float64).pin_memory()
        h2d = sub.coords.nbytes + sub.conn.nbytes + dd.nbytes  # what gl_mesh_upload copies: coordinates + 64-bit point ids (usize)
        d2h = n_e2e * 576 * 8

        def e2e_step():
            dm = ctx.upload(sub)
            integ.mat_10_bdb(ctx, dm, None, cargs, coef, out=host_out)
            dm.free()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e2e * e2e_steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "workload": f"{nx}x{ny}x{layers} cells per GPU per step ({n_e2e} cells): gl_mesh_upload (pinned host coords + "
                           f"connectivity) + gl_mat_10_bdb into a pinned host buffer; PCIe D2H of the matrices dominates",
               "parity_first_cell_rel": float(np.max(np.abs(host_out[:576].numpy() - out[:576].cpu().numpy())) /
                                              np.max(np.abs(k0)))}
        del host_out

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        kernel_ms = ms / args.steps  # rank 0's own kernel time per launch (one kernel per step)
        achieved = BYTES_PER_CELL * ncell / (kernel_ms * 1e-3) / 1e9
        traffic_pc = ncu_traffic_per_cell()
        fp64_peak = ctx.fp64_peak("dfma", 8192)
        dmma_peak = ctx.fp64_peak("dmma", 2048)
        fp64_ach = FLOPS_PER_CELL * ncell / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": None if traffic_pc is None else traffic_pc * ncell, "peak_source": peak_src,
            "kernel": "hex8 mat_10_bdb", "kernel_ms": kernel_ms, "algorithmic_bytes_per_cell": BYTES_PER_CELL,
            "fp64": {"achieved": fp64_ach, "peak_dfma_measured": fp64_peak, "peak_dmma_measured": dmma_peak,
                     "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak, "algorithmic_flops_per_cell": FLOPS_PER_CELL,
                     "note": "Hex8 stiffness sits on the roofline ridge (8.0 flop/B): both fractions are reported"},
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import pyoracle as po

            threads = po.ncores()
            rate1, n1, _ = cpu_leg(args, sub, dd, 20000, 1)
            sample = args.cpu_sample or int(min(sub.ncell, max(20000, rate1 * threads * 8.0)))
            rate, n, dt = cpu_leg(args, sub, dd, sample, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {n} cells of the same mesh (its first {layers} z-layers), {dt:.1f} s, OpenMP over cells with {threads} threads "
                             f"(1 thread, the reference's own execution model: {rate1:.0f} elements/s)",
                   "single_thread_value": rate1}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "checks": checks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
