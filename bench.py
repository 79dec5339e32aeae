#!/usr/bin/env python3
"""bench.py -- Mrays/s of the volume-traversal path on synthetic volumes of the BASELINE shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--traversal NAME] [--gather ipc|nccl]

A "step" is one frame: one pass of the traversal kernel over every pixel of the frame, with
the camera of the workload's script (frame i of the script on step i).  Timed region = K
frames back to back with the volume already resident in HBM (the reference's own benchmark
recipe: --discard-output, README.md:85-89), device-timed, max over ranks.  `e2e` = the same
frames through the public C-ABI call with the camera coming from host memory and the finished
RGBA8 frame copied back to pinned host memory every step.

N > 1 (torchrun, one process per GPU): the frame is split by screen region across the GPUs --
16-row stripes dealt round-robin, the partition a headless.conf with many `device {}` blocks
expresses -- with the volume replicated, and every GPU's kernel stores its finished pixels
straight into rank 0's frame buffer over NVLink (CUDA IPC peer memory; --gather nccl uses an
NCCL gather of contiguous bands instead).  Frame size grows with N (rays per GPU fixed = weak).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 1729
EMISSION = 10.0  # -e 10 as in the reference's README

# BASELINE.md section 4
WORKLOADS = {
    "cfg1": dict(volume=("bunny", 512, 361, 512), frame=(1920, 1080), camera="camera-single", traversal="dda",
                 desc="DDA, V-bunny 512x361x512 RGBA8 grid, 1920x1080, camera-single"),
    "cfg2": dict(volume=("bunny", 512, 361, 512), frame=(1920, 1080), camera="camera", traversal="esvo",
                 desc="ESVO, V-bunny 512x361x512 as lossless SVO (convert --chan-diff 0), 1920x1080, camera.txt path"),
    "cfg3": dict(volume=("tng", 1024, 1024, 1024), frame=(3840, 2160), camera="camera", traversal="dda",
                 desc="DDA, V-tng 1024^3 RGBA8 grid (4 GiB), 3840x2160, camera.txt path"),
    "cfg3r": dict(volume=("tng", 1024, 1024, 1024), frame=(3840, 2160), camera="camera", traversal="svo-rope",
                  desc="svo-rope, V-tng 1024^3 as lossless rope SVO (GPU convert --chan-diff 0 --rope), 3840x2160, camera.txt path"),
    "cfg4": dict(volume=("tng", 2048, 2048, 2048), frame=(3840, 2160), camera="camera", traversal="dda",
                 desc="DDA, V-tng 2048^3 RGBA8 grid (32 GiB), 3840x2160, camera.txt path"),
    "cfg4e": dict(volume=("tng", 2048, 2048, 2048), frame=(3840, 2160), camera="camera", traversal="esvo",
                  desc="ESVO, V-tng 2048^3 as lossless SVO (GPU convert --chan-diff 0), 3840x2160, camera.txt path"),
    "cfg5": dict(volume=("tng", 1024, 1024, 1024), frame=(7680, 4320), camera="camera-rotate", traversal="dda",
                 desc="DDA, V-tng 1024^3, 7680x4320 split by screen region, camera-rotate"),
}


_JSON_OUT = sys.stdout


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def frame_for(base, n_gpus, weak):
    from xenodon_b200 import distributed as xd
    return xd.frame_for(base, n_gpus, weak)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="xn_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(workload, kernel_name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(f"{workload}/{kernel_name}")
        return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"]) if e else None
    except Exception:
        return None


def load_cameras(name):
    from xenodon_b200 import cameras
    return cameras.SCRIPTS[name]()


def cam_tuple(frames, i):
    f = frames[i % len(frames)]
    return (tuple(f[0]), tuple(f[1]), tuple(f[2]))


# --------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (the reference's Vulkan build cannot run here)
# --------------------------------------------------------------------------------------------
def oracle_sample_rows(h, bands=8, rows_per_band=8):
    """Bounded sample of a frame: `bands` groups of rows spread evenly over the frame height."""
    out = []
    for b in range(bands):
        y0 = int((b + 0.5) * h / bands) - rows_per_band // 2
        out.append((max(0, y0), rows_per_band))
    return out


def run_oracle_frames(workload, traversal, frame, cams, steps, warmup, host_grid, tree, prefer_ref=False):
    """Times the CPU implementation on a bounded sample of each frame; returns (Mrays/s, info).

    prefer_ref: use oracle/_ref/libxnref_glsl.so -- the reference's OWN shader text compiled as
    C++ (built where the reference checkout existed; it travels with the repo) -- when present
    (kind "reference"); otherwise the C restatement oracle/xn_oracle.c (kind "port")."""
    from oracle import xo, xref
    use_ref = prefer_ref and xref.available()
    W, H = frame
    bands = oracle_sample_rows(H)
    threads = os.cpu_count() or 1
    rays = 0
    t_total = 0.0
    for i in range(warmup + steps):
        cam = cam_tuple(cams, i)
        t0 = time.perf_counter()
        for (y0, rows) in bands:
            kw = dict(camera=cam, output=(0, y0, W, rows), display=(0, 0, W, H), emission=EMISSION)
            if use_ref:
                if traversal == "dda":
                    xref.render("dda", grid=host_grid, **kw)
                else:
                    xref.render(traversal, nodes=tree.nodes, side=tree.side, **kw)
            elif traversal == "dda":
                xo.render("dda", grid=host_grid, threads=threads, want_stats=False, **kw)
            else:
                xo.render(traversal, nodes=tree.nodes, side=tree.side, threads=threads, want_stats=False, **kw)
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_total += dt
            rays += sum(r for _, r in bands) * W
    mrays = rays / t_total / 1e6
    sample = (f"{steps} frames x {len(bands)} bands of {bands[0][1]} rows ({sum(r for _, r in bands)}/{H} rows "
              f"of each {W}x{H} frame), all {threads} host threads")
    return mrays, dict(cores=threads, sample=sample, seconds=t_total, rays=rays,
                       kind="reference" if use_ref else "port")


# --------------------------------------------------------------------------------------------
def main():
    # keep stdout for the ONE JSON line: libraries (NCCL's version banner, torchrun notices)
    # that print to fd 1 are sent to stderr instead
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="xenodon_b200", choices=["xenodon_b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--traversal", default=None)
    ap.add_argument("--gather", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--strong", action="store_true", help="keep the frame size fixed as N grows")
    ap.add_argument("--no-extras", action="store_true", help="skip per-traversal extras and the CPU baseline")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    n_gpus = world if world > 1 else 1
    if args.gpus > 1 and world == 1:
        log("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
        sys.exit(2)

    wl = WORKLOADS[args.workload]
    traversal = args.traversal or wl["traversal"]
    kind_name, nx, ny, nz = wl["volume"]
    weak = not args.strong
    W, H = frame_for(wl["frame"], n_gpus, weak)
    cams = load_cameras(wl["camera"])
    steps, warmup = args.steps, max(args.warmup, 0)

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference_arm(args, wl, traversal, (W, H), cams, n_gpus)
        return

    import torch
    import torch.distributed as dist

    import xenodon_b200 as xb

    torch.cuda.set_device(local_rank)
    if n_gpus > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind = xb.SYNTH_BUNNY if kind_name == "bunny" else xb.SYNTH_TNG

    ctx = xb.Context(local_rank)
    t_setup = time.perf_counter()
    ctx.synth_grid(kind, nx, ny, nz, SEED)  # generated directly in HBM
    host_grid = None
    tree = None
    need_tree = traversal != "dda"
    extras = [] if args.no_extras or n_gpus > 1 or args.workload != "cfg2" else ["dda", "svo-rope", "svo-df", "svo-naive"]
    tree_side = None
    want_host = (n_gpus == 1 and not args.no_extras and rank == 0)  # the CPU baseline needs host copies
    if need_tree or extras:
        rope = traversal == "svo-rope" or "svo-rope" in extras
        # `xenodon convert --chan-diff 0 [--rope]` on the GPU, from the resident grid (byte-identical
        # to the host builder); the node array comes back to the host only for the CPU baseline
        tree, bstats, n_nodes, tree_side = ctx.convert_resident_grid(
            chan_diff=0, type=xb.TYPE_ROPE if rope else xb.TYPE_SPARSE, bind=True, want_nodes=want_host)
        if rank == 0:
            log(f"[bench] SVO (GPU convert --chan-diff 0{' --rope' if rope else ''}): {n_nodes} nodes, side "
                f"{tree_side}, depth {bstats['depth']}, {n_nodes * 64 / 2**20:.0f} MiB resident")
    if rank == 0:
        log(f"[bench] setup {time.perf_counter() - t_setup:.1f} s; frame {W}x{H}, traversal {traversal}, N={n_gpus}")

    grid_layout = {xb.LAYOUT_LINEAR: "x-major linear", xb.LAYOUT_BRICKED: "8x8x8 bricks, Morton inside",
                   xb.LAYOUT_TEXTURE: "3-D CUDA array (block-linear), texture units"}.get(
        ctx.grid_layout()[0], "none")
    display = (0, 0, W, H)
    ctx.set_params((1, 1, 1), (nx, ny, nz) if traversal == "dda" else (tree_side,) * 3, EMISSION)

    # ---- partition + gather plumbing ----
    frame_ptr = None
    band = None
    if n_gpus == 1:
        ctx.set_target(display, display)
    elif args.gather == "ipc":
        # every rank shades its stripes of the full frame and stores them into rank 0's frame
        ctx.set_target(display, display)
        ctx.set_interleave(n_gpus, rank)
        # two frames: while rank 0 copies frame i to the host, frame i+1 is stored into the other
        if rank == 0:
            made = [ctx.frame_buffer_create(W, H) for _ in range(2)]
            frame_ptrs = [m[0] for m in made]
            obj = [[m[1] for m in made]]
        else:
            obj = [None]
        dist.broadcast_object_list(obj, src=0)
        if rank != 0:
            frame_ptrs = [ctx.frame_buffer_open(h) for h in obj[0]]
        frame_ptr = frame_ptrs[0]
        ctx.set_target_buffer(frame_ptr, W)
    else:
        from xenodon_b200 import distributed as xd
        rows = xd.band_rows(H, n_gpus)
        band = (0, rows[rank], W, rows[rank + 1] - rows[rank])
        ctx.set_target(band, display)
    my_rays = ctx.owned_rays()

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_nccl():
        """NCCL gather of contiguous bands to rank 0 (the plain-library baseline of the gather)."""
        tile = torch.empty((band[3], W), dtype=torch.int32, device="cuda")
        ctx.set_target_buffer(tile.data_ptr(), W)
        return tile

    tile = gather_nccl() if (n_gpus > 1 and args.gather == "nccl") else None
    gathered = None
    if tile is not None and rank == 0:
        sizes = [rows[r + 1] - rows[r] for r in range(n_gpus)]
        gathered = [torch.empty((s, W), dtype=torch.int32, device="cuda") for s in sizes]

    def do_gather():
        if tile is None:
            return
        # bands have different heights: point-to-point (grouped) instead of dist.gather
        if rank == 0:
            reqs = [dist.irecv(gathered[r], src=r) for r in range(1, n_gpus)]
            gathered[0].copy_(tile)
            for q in reqs:
                q.wait()
        else:
            dist.send(tile, dst=0)

    # ---- kernel-only timed region ----
    for i in range(warmup):
        ctx.render(traversal, cam_tuple(cams, i))
        ctx.sync()
        do_gather()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    kernel_ms = []
    ctx.mark(0)
    for i in range(steps):
        ctx.render(traversal, cam_tuple(cams, warmup + i))
        if tile is not None:
            ctx.sync()
            do_gather()
    ctx.mark(1)
    region_ms = ctx.mark_elapsed()
    barrier()
    launches = ctx.launch_count() - launches0

    # per-frame kernel durations (CUDA events on the launching stream), for the roofline
    for i in range(steps):
        ctx.render(traversal, cam_tuple(cams, warmup + i))
        kernel_ms.append(ctx.sync())
    barrier()

    # ---- end-to-end region: camera from the host every step, frame back to pinned host memory ----
    pinned = [xb.PinnedFrame(W, H), xb.PinnedFrame(W, H)] if rank == 0 else None
    e2e_d2h = W * H * 4
    for i in range(min(warmup, 3)):
        if n_gpus == 1:
            ctx.render_download_async(traversal, cam_tuple(cams, i), pinned[i & 1])
            ctx.sync()
    barrier()
    t0 = time.perf_counter()
    if n_gpus == 1:
        for i in range(steps):
            ctx.render_download_async(traversal, cam_tuple(cams, warmup + i), pinned[i & 1])
        ctx.sync()
    else:
        if tile is None:
            ctx.set_target_buffer(frame_ptrs[0], W)
        for i in range(steps):
            ctx.render(traversal, cam_tuple(cams, warmup + i))
            ctx.sync()
            if tile is not None:
                do_gather()
                torch.cuda.synchronize()
                if rank == 0:
                    y = 0
                    dst = torch.from_numpy(pinned[i & 1].array.view(np.int32).reshape(H, W))
                    for g in gathered:
                        dst[y:y + g.shape[0]].copy_(g)
                        y += g.shape[0]
            else:
                if rank == 0:
                    ctx.copy_sync()  # frame i-1 is on the host before anyone may overwrite its buffer
                dist.barrier()  # every rank's peer stores of frame i have landed in rank 0's frame
                if rank == 0:
                    ctx.frame_buffer_read_async(frame_ptrs[i & 1], W, H, pinned[i & 1])
                if i + 1 < steps:
                    ctx.set_target_buffer(frame_ptrs[(i + 1) & 1], W)
        if rank == 0 and tile is None:
            ctx.copy_sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    last_frame = pinned[(steps - 1) & 1].array.copy() if rank == 0 else None

    clocks = sampler.stop() if sampler else None

    # ---- reduce over ranks: max time, sum rays ----
    region_all, e2e_all, rays_all = region_ms, e2e_s, my_rays
    if n_gpus > 1:
        t = torch.tensor([region_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_all, e2e_all = t.tolist()
        r = torch.tensor([my_rays], dtype=torch.int64, device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays_all = int(r.item())

    # ---- algorithmic bytes (instrumented pass, untimed) for the roofline of the dominant kernel ----
    alg_bytes = 0
    alg_steps = 0
    stat_frames = list(range(0, steps, max(1, steps // 30)))  # every ~5th frame; scaled to all frames
    for i in stat_frames:
        _, _, (s, b) = ctx.stats_pass(traversal, cam_tuple(cams, warmup + i), per_ray=False)
        alg_steps += s
        alg_bytes += b
    scale = steps / len(stat_frames)
    alg_bytes_per_frame = (alg_bytes * scale + my_rays * 4.0 * steps) / steps  # + 4 B pixel store per ray
    alg_steps_per_frame = alg_steps * scale / steps

    if rank != 0:
        if n_gpus > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = rays_all * steps / (region_all / 1e3) / 1e6
    e2e_value = rays_all * steps / e2e_all / 1e6
    peak, peak_kind = measured_peaks()
    mean_kernel_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes_per_frame / (mean_kernel_ms / 1e3) / 1e9
    kernel_name = {"dda": "dda_kernel", "esvo": "esvo_kernel", "svo-rope": "svo_rope_kernel",
                   "svo-df": "svo_df_kernel", "svo-naive": "svo_naive_kernel"}[traversal]
    if traversal == "dda" and ctx.grid_layout()[0] == xb.LAYOUT_TEXTURE:
        kernel_name = "dda_tex_kernel"
    result = {
        "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(region_all / steps, 5), "higher_is_better": True,
        "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": round(steps / (region_all / 1e3), 2),
        "config": {
            "workload": f"{args.workload}: {wl['desc']}", "traversal": traversal, "frame": f"{W}x{H}",
            "volume": f"{kind_name} {nx}x{ny}x{nz} seed {SEED}", "camera": wl["camera"], "emission": EMISSION,
            "grid_layout": grid_layout if traversal == "dda" else None,
            "partition": ("single region" if n_gpus == 1 else
                          f"16-row stripes round-robin over {n_gpus} GPUs, peer stores into rank 0's frame (CUDA IPC/NVLink)"
                          if args.gather == "ipc" else f"{n_gpus} horizontal bands, NCCL send/recv gather to rank 0"),
            "l2": "inputs larger than L2 (volume resident in HBM exceeds 126 MB); camera changes every step",
        },
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 36 * n_gpus,
                "d2h_bytes_per_step": e2e_d2h, "frames_per_s": round(steps / e2e_all, 2)},
        "gpu_launches": int(launches) * n_gpus,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": round(achieved, 1), "peak": peak,
                     "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": measured_traffic(args.workload, kernel_name) if n_gpus == 1 else None,
                     "traffic_source": "profiles/traffic.json (ncu --set full capture of one camera frame)",
                     "algorithmic_bytes_per_launch": round(alg_bytes_per_frame),
                     "steps_per_launch": round(alg_steps_per_frame),
                     "mean_kernel_ms": round(mean_kernel_ms, 5),
                     "note": "requested bytes (4 B per texel fetch / node-field read as the shader writes them "
                             "+ 4 B per pixel); cache hits make this exceed DRAM traffic"},
        "reference_stats": {"mray/s (rays / summed device ms)": round(
            rays_all * steps / n_gpus / (sum(kernel_ms)) / 1e3, 2)},
    }

    # ---- extras on the same volume (N=1, cfg2 only): every traversal, same frames ----
    per_traversal = {traversal: round(value, 2)}
    for t in extras:
        if t == traversal:
            continue
        ctx.set_params((1, 1, 1), (nx, ny, nz) if t == "dda" else (tree_side,) * 3, EMISSION)
        for i in range(3):
            ctx.render(t, cam_tuple(cams, i))
        ctx.sync()
        ctx.mark(0)
        for i in range(steps):
            ctx.render(t, cam_tuple(cams, warmup + i))
        ctx.mark(1)
        ms = ctx.mark_elapsed()
        per_traversal[t] = round(W * H * steps / (ms / 1e3) / 1e6, 2)
    if len(per_traversal) > 1:
        result["per_traversal_mrays_s"] = per_traversal

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample ----
    if n_gpus == 1 and not args.no_extras:
        try:
            if traversal == "dda":
                if nx * ny * nz * 4 > (8 << 30):
                    raise RuntimeError("volume too large for the host-side baseline sample")
                host_grid = ctx.download_grid()
            else:
                host_grid = xb.Grid(np.zeros((1, 1, 1, 4), np.uint8))  # unused by the octree traversals
            sample_steps = max(1, min(steps, 12))
            sub = cams[np.linspace(0, len(cams) - 1, sample_steps).astype(int)] if len(cams) > 1 else cams
            mr, info = run_oracle_frames(args.workload, traversal, (W, H), sub, sample_steps, 1,
                                         host_grid.data, tree)
            result["cpu_baseline"] = {"value": round(mr, 3), "unit": "Mrays/s", "cores": info["cores"],
                                      "kind": info["kind"], "sample": info["sample"]}
            # parity spot check of the last e2e frame against the oracle (not timed)
            from oracle import xo
            cam = cam_tuple(cams, warmup + steps - 1)
            y0 = H // 2 - 8
            kw = dict(camera=cam, output=(0, y0, W, 16), display=display, emission=EMISSION, want_stats=False)
            ref = (xo.render("dda", grid=host_grid.data, **kw) if traversal == "dda"
                   else xo.render(traversal, nodes=tree.nodes, side=tree.side, **kw))[0]
            d = np.abs(ref.astype(int) - last_frame[y0:y0 + 16].astype(int)).max(axis=-1)
            result["parity_check"] = {"rows": 16, "within_1_of_255": float((d <= 1).mean()), "max_diff": int(d.max())}
        except Exception as e:  # the baseline must never take the GPU number down with it
            result["cpu_baseline"] = {"error": str(e)}

    print(json.dumps(result), file=_JSON_OUT, flush=True)
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference_arm(args, wl, traversal, frame, cams, n_gpus):
    """--impl reference: the reference's CPU-runnable stand-in (the oracle port) on the host cores.
    The volume comes from the host generator (bit-identical to the device generator)."""
    import xenodon_b200 as xb  # host-side formats / generators only; no GPU call on this arm
    kind_name, nx, ny, nz = wl["volume"]
    W, H = frame
    steps, warmup = args.steps, max(args.warmup, 0)
    if nx * ny * nz > 1100 ** 3:
        print(json.dumps({"impl": "reference", "unavailable": "host copy of this volume exceeds the arm's budget"}),
              file=_JSON_OUT, flush=True)
        return
    log(f"[reference arm] generating {kind_name} {nx}x{ny}x{nz} on the host ...")
    grid = xb.Grid.synthetic(xb.SYNTH_BUNNY if kind_name == "bunny" else xb.SYNTH_TNG, nx, ny, nz, SEED)
    tree = None
    if traversal != "dda":
        log("[reference arm] convert --chan-diff 0 ...")
        tree, _ = xb.build_octree(grid, chan_diff=0, type=xb.TYPE_ROPE if traversal == "svo-rope" else xb.TYPE_SPARSE)
    # bounded: at most ~40 sampled frames spread over the requested steps
    n = max(1, min(steps, 40))
    idx = np.linspace(warmup, warmup + steps - 1, n).astype(int)
    sub = cams[[i % len(cams) for i in idx]]
    t0 = time.perf_counter()
    mr, info = run_oracle_frames(args.workload, traversal, (W, H), sub, n, min(warmup, 1), grid.data, tree,
                                 prefer_ref=True)
    wall = time.perf_counter() - t0
    rows = sum(r for _, r in oracle_sample_rows(H))
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": round(mr, 3), "unit": "Mrays/s", "n_gpus": n_gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(info["seconds"] / n * 1e3 * (H / rows), 3),
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "traversal": traversal, "frame": f"{W}x{H}",
                   "volume": f"{kind_name} {nx}x{ny}x{nz} seed {SEED}", "camera": wl["camera"],
                   "emission": EMISSION,
                   "note": "the reference's Vulkan build cannot run in this image (no loader/ICD/glslc); this arm "
                           "times its compute-shader text compiled as C++ (oracle/_ref, kind 'reference') or, where "
                           "that library is absent, the C restatement oracle/xn_oracle.c (kind 'port')"},
        "cpu_baseline": {"value": round(mr, 3), "unit": "Mrays/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": round(mr, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 1),
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


if __name__ == "__main__":
    main()
