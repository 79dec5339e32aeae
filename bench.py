#!/usr/bin/env python3
"""bench.py -- Mrays/s of the volume-traversal path on synthetic volumes of the BASELINE shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload cfg1|cfg2|cfg3|cfg3r|cfg4|cfg4e|cfg5] [--traversal NAME]
                    [--gather host|ipc|nccl] [--weak] [--no-extras]

A "step" is one frame: one pass of the traversal kernel over every pixel of the frame.  The K
steps are spread over the WHOLE camera script of the workload (frame floor(i * len / K) on step
i, the reference's own benchmark recipe runs the whole file: README.md:85-89), so that exterior,
fly-over and interior views are all sampled whatever K is.

Defaults (BASELINE.json):
  --gpus 1   cfg4: DDA through the 2048^3 RGBA8 grid (32 GiB, beyond the reference's 4 GB binding
             limit) at 3840x2160 over camera.txt -- the largest single-GPU configuration.  The same
             line carries `per_config`: cfg1, cfg2 (every traversal), cfg3, cfg3r and cfg4e measured
             the same way in the same process.
  --gpus N   cfg5 STRONG: one 7680x4320 frame of the 1024^3 volume split by screen region over
             the N GPUs (camera-rotate.txt), volume replicated.

`value` = rays of the K frames / device time of the K back-to-back launches with the volume
resident in HBM (the reference's --discard-output recipe), max over ranks.  `e2e` = the same
frames through the public C-ABI call with the camera coming from host memory and every finished
RGBA8 frame delivered to page-locked host memory, wall-clocked between barriers.

N > 1 (torchrun, one process per GPU): 16-row stripes dealt round-robin over the ranks (the
partition a headless.conf with many `device {}` blocks expresses).  In the kernel-timed region every
rank's kernel stores its finished pixels straight into rank 0's frame over NVLink (CUDA IPC peer
memory).  In the e2e region (--gather host, default) every rank copies its own stripes into ONE
page-locked host frame shared by all ranks (POSIX shared memory registered with CUDA in every
process), so the frame leaves over N PCIe links instead of rank 0's one; completion is signalled
per rank by a stream-ordered flag in the same shared memory, no collective and no host
synchronisation inside the loop.  --gather ipc / nccl keep the NVLink-gather-then-one-link and the
NCCL baselines.
"""
from __future__ import annotations

import argparse
import json
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
                 desc="ESVO, V-bunny 512x361x512 as lossless SVO (convert --chan-diff 0), 1920x1080, whole camera.txt path"),
    "cfg3": dict(volume=("tng", 1024, 1024, 1024), frame=(3840, 2160), camera="camera", traversal="dda",
                 desc="DDA, V-tng 1024^3 RGBA8 grid (4 GiB), 3840x2160, whole camera.txt path"),
    "cfg3r": dict(volume=("tng", 1024, 1024, 1024), frame=(3840, 2160), camera="camera", traversal="svo-rope",
                  desc="svo-rope, V-tng 1024^3 as lossless rope SVO (GPU convert --chan-diff 0 --rope), 3840x2160, whole camera.txt path"),
    "cfg4": dict(volume=("tng", 2048, 2048, 2048), frame=(3840, 2160), camera="camera", traversal="dda",
                 desc="DDA, V-tng 2048^3 RGBA8 grid (32 GiB), 3840x2160, whole camera.txt path"),
    "cfg4e": dict(volume=("tng", 2048, 2048, 2048), frame=(3840, 2160), camera="camera", traversal="esvo",
                  desc="ESVO, V-tng 2048^3 as lossless SVO (GPU convert --chan-diff 0), 3840x2160, whole camera.txt path"),
    "cfg5": dict(volume=("tng", 1024, 1024, 1024), frame=(7680, 4320), camera="camera-rotate", traversal="dda",
                 desc="DDA, V-tng 1024^3, 7680x4320 split by screen region, whole camera-rotate.txt path"),
}

KERNEL_NAMES = {"dda": "dda_kernel", "esvo": "esvo_kernel", "svo-rope": "svo_rope_kernel",
                "svo-df": "svo_df_kernel", "svo-naive": "svo_naive_kernel"}

_JSON_OUT = sys.stdout


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def frame_for(base, n_gpus, weak):
    from xenodon_b200 import distributed as xd
    return xd.frame_for(base, n_gpus, weak)


def frame_schedule(n_script: int, count: int):
    """Script frame of each of `count` steps: spread evenly over the whole script."""
    return [(i * n_script) // max(count, 1) for i in range(count)]


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
            # median over the samples taken while the GPU was clocked up (under load)
            busy = [s for s in sm if s >= 0.5 * max(mx)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profile_entry(workload, kernel_name):
    """What the committed ncu capture of this workload's dominant kernel says (profiles/traffic.json):
    DRAM bytes per launch and which unit bounds the kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"{workload}/{kernel_name}")
    except Exception:
        return None


def load_cameras(name):
    from xenodon_b200 import cameras
    return cameras.SCRIPTS[name]()


def cam_tuple(frames, i):
    f = frames[i % len(frames)]
    return (tuple(f[0]), tuple(f[1]), tuple(f[2]))


def host_threads() -> int:
    """Host cores this process may use (torchrun's OMP_NUM_THREADS=1 does not apply to the CPU arms:
    they size their own thread pools)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def set_openmp_threads(n: int) -> int:
    """The reference shader library (oracle/_ref) is `#pragma omp parallel for`: make its thread
    count explicit instead of inheriting OMP_NUM_THREADS from the launcher.  Returns what is in force."""
    import ctypes
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_dynamic(0)
        gomp.omp_set_num_threads(n)
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


# --------------------------------------------------------------------------------------------
# CPU arms: the reference's own shader text compiled as C++ (oracle/_ref, kind "reference") or the
# C restatement (oracle/xn_oracle.c, kind "port").  The reference's Vulkan build cannot run here
# (no loader / ICD / glslc).
# --------------------------------------------------------------------------------------------
def oracle_sample_rows(h, bands=8, rows_per_band=8):
    """Bounded sample of a frame: `bands` groups of rows spread evenly over the frame height."""
    out = []
    for b in range(bands):
        y0 = int((b + 0.5) * h / bands) - rows_per_band // 2
        out.append((max(0, y0), rows_per_band))
    return out


def run_oracle_frames(traversal, frame, cam_list, host_grid, tree, prefer_ref=False, warm=1, bands=None):
    """Times the CPU implementation on a bounded row sample of each camera in cam_list (the first
    `warm` are untimed); returns (Mrays/s, info)."""
    from oracle import xo, xref
    use_ref = prefer_ref and xref.available()
    W, H = frame
    bands = bands or oracle_sample_rows(H)
    threads = host_threads()
    in_force = set_openmp_threads(threads) if use_ref else threads
    rays = 0
    t_total = 0.0
    for i, cam in enumerate(cam_list):
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
        if i >= warm:
            t_total += dt
            rays += sum(r for _, r in bands) * W
    n_timed = max(0, len(cam_list) - warm)
    mrays = rays / t_total / 1e6 if t_total > 0 else 0.0
    sample = (f"{n_timed} frames spread over the script x {len(bands)} bands of {bands[0][1]} rows "
              f"({sum(r for _, r in bands)}/{H} rows of each {W}x{H} frame), {in_force} host threads"
              f"{' (OpenMP, set by the arm itself)' if use_ref else ''}")
    return mrays, dict(cores=in_force, sample=sample, seconds=t_total, rays=rays, frames=n_timed,
                       kind="reference" if use_ref else "port")


# --------------------------------------------------------------------------------------------
# GPU arm helpers
# --------------------------------------------------------------------------------------------
def timed_frames(ctx, traversal, cams, sched, warm_sched):
    """Device time (ms) of len(sched) back-to-back launches, after the warm-up launches."""
    for f in warm_sched:
        ctx.render(traversal, cam_tuple(cams, f))
    ctx.sync()
    ctx.mark(0)
    for f in sched:
        ctx.render(traversal, cam_tuple(cams, f))
    ctx.mark(1)
    return ctx.mark_elapsed()


def per_launch_ms(ctx, traversal, cams, sched):
    out = []
    for f in sched:
        ctx.render(traversal, cam_tuple(cams, f))
        out.append(ctx.sync())
    return out


def algorithmic_bytes(ctx, traversal, cams, sched, rays_per_frame, max_frames=12):
    """Mean algorithmic bytes and steps per launch (instrumented STATS pass, untimed): 4 B per texel
    fetch / node-field read as the shader source writes them + 4 B pixel store per ray."""
    pick = sched[::max(1, len(sched) // max_frames)] or sched
    tot_s = tot_b = 0
    for f in pick:
        _, _, (s, b) = ctx.stats_pass(traversal, cam_tuple(cams, f), per_ray=False)
        tot_s += s
        tot_b += b
    return tot_b / len(pick) + 4.0 * rays_per_frame, tot_s / len(pick), len(pick)


def roofline_block(workload, traversal, kernel_name, alg_bytes, alg_steps, n_stat, mean_ms, n_gpus):
    peak, peak_kind = measured_peaks()
    achieved = alg_bytes / (mean_ms / 1e3) / 1e9
    prof = profile_entry(workload, kernel_name) if n_gpus == 1 else None
    traffic = (int(prof["dram_bytes_read"]) + int(prof["dram_bytes_write"])) if prof else None
    block = {
        # which unit the committed ncu capture shows saturated: "hbm" (DRAM), "tex" (texture unit
        # wavefronts), "issue" (warp instruction issue at the measured lanes per instruction)
        "bound": (prof or {}).get("bound", "hbm"),
        "kernel": kernel_name, "achieved": round(achieved, 1), "peak": peak, "peak_kind": peak_kind,
        "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
        "traffic_source": "profiles/traffic.json (ncu --set full capture of one camera frame)" if prof else None,
        "algorithmic_bytes_per_launch": round(alg_bytes), "steps_per_launch": round(alg_steps),
        "mean_kernel_ms": round(mean_ms, 5), "stat_frames": n_stat,
        "note": "achieved = requested bytes (4 B per texel fetch / node-field read as the shader writes them + 4 B "
                "per pixel) / kernel time; cache hits and fetches the skip table makes unnecessary mean the DRAM "
                "traffic is far below it -- see dram_frac and the ncu evidence",
    }
    if prof:
        for k in ("dram_frac_of_measured_peak", "issue_active", "lanes_per_instruction", "l2_hit", "l2_gb_per_s",
                  "capture_ms", "capture", "evidence"):
            if k in prof:
                block[k] = prof[k]
        if traffic is not None and "capture_ms" in prof:
            block["dram_frac"] = round(traffic / (prof["capture_ms"] / 1e3) / 1e9 / peak, 4)
        # how close the unit named by `bound` is to its own ceiling in that capture
        near = {"issue": max(prof.get("issue_active", 0.0), prof.get("pipe_alu", 0.0), prof.get("pipe_fma", 0.0)),
                "tex": prof.get("tex_wavefronts", 0.0), "lsu": prof.get("lsu_wavefronts", 0.0),
                "hbm": 100.0 * prof.get("dram_frac_of_measured_peak", 0.0)}.get(block["bound"])
        if near is not None:
            block["bound_frac"] = round(near / 100.0, 3)
    return block


def compulsory_traffic(xb, ctx, traversal, cams, sched, kernel_ms):
    """SURVEY 8(d)'s sector-granular lower bound for the DDA, on the schedule's 4/5-point frame (script
    frame 120 of 150 = the frame the committed ncu capture holds): distinct voxels and distinct 32-byte
    sectors (of the x-major linear volume) one frame fetches -- as the shader requests them, and as the
    skip table leaves them -- from the instrumented pass (untimed)."""
    if traversal != "dda" or ctx.grid_layout()[0] != xb.LAYOUT_TEXTURE:
        return None
    try:
        i = min(len(sched) - 1, (4 * len(sched)) // 5)
        cam = cam_tuple(cams, sched[i])
        out = {"script_frame": int(sched[i]), "kernel_ms_of_that_frame": round(float(kernel_ms[i]), 4)}
        for key, skip in (("as_requested_by_the_shader", False), ("fetched_with_the_skip_table", True)):
            if skip and os.environ.get("XN_DDA_SKIP", "1") == "0":
                continue
            vox, sec = ctx.touch_pass(cam, use_skip_table=skip)
            out[key] = {"distinct_voxels": vox, "distinct_sectors_32B": sec, "bytes": sec * 32,
                        "gb_per_s_at_that_frame": round(sec * 32 / (float(kernel_ms[i]) / 1e3) / 1e9, 1)}
        out["note"] = ("a frame cannot move fewer bytes from HBM than the distinct sectors it fetches (x 32 B; "
                       "linear-layout sectors, the texture residency's tiles group voxels differently); compare "
                       "with roofline.traffic (ncu DRAM bytes of the same frame)")
        return out
    except Exception as e:  # an explanatory figure must not take the bench line down
        return {"error": str(e)}


def kernel_name_for(xb, ctx, traversal):
    name = KERNEL_NAMES[traversal]
    if traversal == "dda" and ctx.grid_layout()[0] == xb.LAYOUT_TEXTURE:
        name = "dda_skip_tex_kernel" if os.environ.get("XN_DDA_SKIP", "1") != "0" else "dda_tex_kernel"
    return name


def measure_config(xb, ctx, name, traversal, frame, cams, steps, warmup, tree_side, dims):
    """One per_config entry: device-timed Mrays/s over the spread schedule + its roofline block."""
    W, H = frame
    ctx.set_target((0, 0, W, H), (0, 0, W, H))
    ctx.set_params((1, 1, 1), dims if traversal == "dda" else (tree_side,) * 3, EMISSION)
    sched = frame_schedule(len(cams), steps)
    warm = frame_schedule(len(cams), max(warmup, 3))
    ms = timed_frames(ctx, traversal, cams, sched, warm)
    kms = per_launch_ms(ctx, traversal, cams, sched)
    alg_b, alg_s, n_stat = algorithmic_bytes(ctx, traversal, cams, sched, W * H)
    value = W * H * len(sched) / (ms / 1e3) / 1e6
    kname = kernel_name_for(xb, ctx, traversal)
    roof = roofline_block(name, traversal, kname, alg_b, alg_s, n_stat, float(np.mean(kms)), 1)
    lb = compulsory_traffic(xb, ctx, traversal, cams, sched, kms)
    if lb:
        roof["sector_lower_bound"] = lb
    return {
        "workload": f"{name}: {WORKLOADS[name]['desc']}" if traversal == WORKLOADS[name]["traversal"]
        else f"{name} volume, --shader {traversal}",
        "traversal": traversal, "value": round(value, 2), "unit": "Mrays/s",
        "ms_per_step": round(ms / len(sched), 5), "frames_per_s": round(len(sched) / (ms / 1e3), 2), "steps": len(sched),
        "roofline": roof,
    }


# --------------------------------------------------------------------------------------------
# shared page-locked host frames for the multi-GPU e2e path
# --------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
def main():
    # keep stdout for the ONE JSON line: libraries (NCCL's version banner, torchrun notices)
    # that print to fd 1 are sent to stderr instead
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="xenodon_b200", choices=["xenodon_b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--traversal", default=None)
    ap.add_argument("--gather", default="host", choices=["host", "ipc", "nccl"])
    ap.add_argument("--weak", action="store_true", help="grow the frame with N (rays per GPU fixed) instead of splitting one frame")
    ap.add_argument("--strong", action="store_true", help="(default) keep the frame size fixed as N grows")
    ap.add_argument("--no-extras", action="store_true", help="skip per_config, per-traversal extras and the CPU baseline")
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

    workload = args.workload or ("cfg4" if n_gpus == 1 else "cfg5")
    wl = WORKLOADS[workload]
    traversal = args.traversal or wl["traversal"]
    weak = bool(args.weak) and not args.strong
    W, H = frame_for(wl["frame"], n_gpus, weak)
    cams = load_cameras(wl["camera"])
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference_arm(args, workload, wl, traversal, (W, H), cams, n_gpus, weak)
        return

    import torch
    import torch.distributed as dist

    import xenodon_b200 as xb
    from xenodon_b200 import distributed as xd

    torch.cuda.set_device(local_rank)
    if n_gpus > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind_name, nx, ny, nz = wl["volume"]
    kind = xb.SYNTH_BUNNY if kind_name == "bunny" else xb.SYNTH_TNG

    ctx = xb.Context(local_rank)
    t_setup = time.perf_counter()
    ctx.synth_grid(kind, nx, ny, nz, SEED)  # generated directly in HBM
    tree = None
    tree_side = None
    extras = not args.no_extras and n_gpus == 1 and rank == 0
    want_host_tree = extras and nx * ny * nz <= 1100 ** 3  # node array back on the host only for the CPU baseline
    if traversal != "dda":
        rope = traversal == "svo-rope"
        # `xenodon convert --chan-diff 0 [--rope]` on the GPU, from the resident grid (byte-identical
        # to the host builder)
        tree, bstats, n_nodes, tree_side = ctx.convert_resident_grid(
            chan_diff=0, type=xb.TYPE_ROPE if rope else xb.TYPE_SPARSE, bind=True, want_nodes=want_host_tree)
        if rank == 0:
            log(f"[bench] SVO (GPU convert --chan-diff 0{' --rope' if rope else ''}): {n_nodes} nodes, side "
                f"{tree_side}, depth {bstats['depth']}")
    if rank == 0:
        log(f"[bench] {workload}: setup {time.perf_counter() - t_setup:.1f} s; frame {W}x{H}, traversal {traversal}, N={n_gpus}")

    grid_layout = {xb.LAYOUT_LINEAR: "x-major linear", xb.LAYOUT_BRICKED: "8x8x8 bricks, Morton inside",
                   xb.LAYOUT_TEXTURE: "3-D CUDA array (block-linear), texture units"}.get(
        ctx.grid_layout()[0], "none")
    display = (0, 0, W, H)
    ctx.set_params((1, 1, 1), (nx, ny, nz) if traversal == "dda" else (tree_side,) * 3, EMISSION)
    sched = frame_schedule(len(cams), steps)
    warm_sched = frame_schedule(len(cams), warmup)

    # ---- partition + gather plumbing ----
    band = None
    frame_ptrs = None
    gather = args.gather if n_gpus > 1 else "single"
    ctx.set_target(display, display)
    if n_gpus > 1 and gather in ("host", "ipc"):
        # every rank shades its 16-row stripes of the full frame; in the kernel-timed region the
        # stripes are stored straight into rank 0's frame over NVLink
        ctx.set_interleave(n_gpus, rank)
        if rank == 0:
            made = [ctx.frame_buffer_create(W, H) for _ in range(2)]
            frame_ptrs = [m[0] for m in made]
            obj = [[m[1] for m in made]]
        else:
            obj = [None]
        dist.broadcast_object_list(obj, src=0)
        if rank != 0:
            frame_ptrs = [ctx.frame_buffer_open(h) for h in obj[0]]
        ctx.set_target_buffer(frame_ptrs[0], W)
    elif n_gpus > 1:
        from xenodon_b200 import distributed as xd
        rows = xd.band_rows(H, n_gpus)
        band = (0, rows[rank], W, rows[rank + 1] - rows[rank])
        ctx.set_target(band, display)
    my_rays = ctx.owned_rays()

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tile = None
    gathered = None
    if gather == "nccl":
        tile = torch.empty((band[3], W), dtype=torch.int32, device="cuda")
        ctx.set_target_buffer(tile.data_ptr(), W)
        if rank == 0:
            sizes = [rows[r + 1] - rows[r] for r in range(n_gpus)]
            gathered = [torch.empty((s, W), dtype=torch.int32, device="cuda") for s in sizes]

    def do_gather():
        """NCCL gather of contiguous bands to rank 0 (the plain-library baseline of the gather)."""
        if tile is None:
            return
        if rank == 0:
            reqs = [dist.irecv(gathered[r], src=r) for r in range(1, n_gpus)]
            gathered[0].copy_(tile)
            for q in reqs:
                q.wait()
        else:
            dist.send(tile, dst=0)

    # ---- kernel-timed region: K frames back to back, finished pixels land in rank 0's HBM ----
    for f in warm_sched:
        ctx.render(traversal, cam_tuple(cams, f))
        ctx.sync()
        do_gather()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    ctx.mark(0)
    for f in sched:
        ctx.render(traversal, cam_tuple(cams, f))
        if tile is not None:
            ctx.sync()
            do_gather()
    ctx.mark(1)
    region_ms = ctx.mark_elapsed()
    barrier()
    launches = ctx.launch_count() - launches0

    # per-frame kernel durations (CUDA events on the launching stream), for the roofline
    kernel_ms = per_launch_ms(ctx, traversal, cams, sched)
    barrier()

    # ---- end-to-end region: camera from the host every step, every frame delivered to page-locked host memory ----
    e2e_d2h = W * H * 4
    shared = None
    pinned = None
    e2e_info = {}
    if n_gpus == 1:
        pinned = [xb.PinnedFrame(W, H), xb.PinnedFrame(W, H)]
        for j, f in enumerate(warm_sched[:3]):
            ctx.render_download_async(traversal, cam_tuple(cams, f), pinned[j & 1])
        ctx.sync()
        barrier()
        t0 = time.perf_counter()
        for i, f in enumerate(sched):
            ctx.render_download_async(traversal, cam_tuple(cams, f), pinned[i & 1])
        ctx.sync()
        barrier()
        e2e_s = time.perf_counter() - t0
        last_frame = pinned[(steps - 1) & 1].array
        e2e_info["path"] = "xn_render_download_async: three device targets in rotation, copy-out of frame i overlaps frames i+1, i+2"
    elif gather == "host":
        # every rank copies its own stripes into the shared page-locked frame (N PCIe links);
        # a stream-ordered flag per rank and slot says "frame seq is complete in this slot"
        ctx.set_target_buffer(None, 0)
        SLOTS = 3
        name = [f"xn_bench_{os.getpid()}_{int(time.time())}" if rank == 0 else None]
        dist.broadcast_object_list(name, src=0)
        if rank == 0:
            shared = xd.SharedHostFrames(name[0], W, H, SLOTS, n_gpus, create=True, register=xb.host_register,
                                           unregister=xb.host_unregister)
        dist.barrier()
        if rank != 0:
            shared = xd.SharedHostFrames(name[0], W, H, SLOTS, n_gpus, create=False, register=xb.host_register,
                                           unregister=xb.host_unregister)
        dist.barrier()
        ring = xd.HostFrameRing(shared, rank)

        def produce(seq, f):
            slot = ring.acquire(seq)  # the consumer has released the frame this slot held
            ctx.render_download_to(traversal, cam_tuple(cams, f), shared.frame_ptr(slot), W)
            ctx.signal_after_copy(shared.flag_ptr(rank, slot), seq)

        def consume(seq):
            ring.wait_complete(seq)  # a consumer (PNG writer, display) would use shared.frames[slot] here
            ring.release(seq)

        base = 0
        for j, f in enumerate(warm_sched[:3]):
            base += 1
            produce(base, f)
            if rank == 0:
                consume(base)
        ctx.sync()
        barrier()
        t0 = time.perf_counter()
        lag = SLOTS - 1  # the consumer trails the producers by this many frames (frames in flight)
        for i, f in enumerate(sched):
            produce(base + 1 + i, f)
            if rank == 0 and i >= lag:
                consume(base + 1 + i - lag)
        if rank == 0:
            for seq in range(max(base + 1, base + 1 + steps - lag), base + steps + 1):
                consume(seq)
        ctx.sync()
        barrier()
        e2e_s = time.perf_counter() - t0
        last_frame = shared.frames[(base + steps) % SLOTS] if rank == 0 else None
        e2e_info["path"] = (f"each rank copies its own 16-row stripes into one page-locked host frame shared by the {n_gpus} "
                            "processes (POSIX shm + cudaHostRegister): the frame leaves over N PCIe links; per-rank "
                            "stream-ordered completion flags, no collective or host sync inside the loop")
        e2e_info["nvlink_bytes_per_step"] = 0
        if rank == 0:
            e2e_info["host_frame_numa"] = shared.numa
        if last_frame is not None:
            last_frame = np.array(last_frame, copy=True)  # the probe below overwrites the segment
        # what the host side of this path can take: every rank copies a device buffer of its share of
        # the frame into its part of the shared segment, all ranks at once, then rank 0 alone
        share = (W * H * 4 // n_gpus) & ~4095
        dev = torch.empty(share, dtype=torch.uint8, device="cuda")
        dst = torch.from_numpy(shared.buf[shared.flag_bytes + rank * share:shared.flag_bytes + (rank + 1) * share])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

        def copy_rate(active):
            barrier()
            ms = 0.0
            if active:
                dst.copy_(dev, non_blocking=True)
                ev[0].record()
                for _ in range(10):
                    dst.copy_(dev, non_blocking=True)
                ev[1].record()
                torch.cuda.synchronize()
                ms = ev[0].elapsed_time(ev[1]) / 10
            t_ = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())

        ms_all = copy_rate(True)
        ms_one = copy_rate(rank == 0)
        e2e_info["d2h_probe"] = {
            "bytes_per_rank": share,
            "all_ranks_at_once_gb_per_s": round(share * n_gpus / (ms_all / 1e3) / 1e9, 1) if ms_all > 0 else None,
            "one_rank_alone_gb_per_s": round(share / (ms_one / 1e3) / 1e9, 1) if ms_one > 0 else None,
            "note": "contiguous device -> shared host frame copies, CUDA-event timed, max over ranks",
        }
        del dev, dst
    else:
        pinned = [xb.PinnedFrame(W, H), xb.PinnedFrame(W, H)] if rank == 0 else None
        if tile is None:
            ctx.set_target_buffer(frame_ptrs[0], W)
        barrier()
        t0 = time.perf_counter()
        for i, f in enumerate(sched):
            ctx.render(traversal, cam_tuple(cams, f))
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
        last_frame = pinned[(steps - 1) & 1].array if rank == 0 else None
        e2e_info["path"] = ("kernel-epilogue peer stores into rank 0's frame over NVLink, then rank 0 copies the whole frame "
                            "over its one PCIe link" if gather == "ipc" else "NCCL send/recv of bands to rank 0, then one PCIe link")
        e2e_info["nvlink_bytes_per_step"] = int(W * H * 4 * (n_gpus - 1) / n_gpus)
    if last_frame is not None:
        last_frame = np.array(last_frame, copy=True)

    clocks = sampler.stop() if sampler else None

    # ---- the same workload on ONE GPU, measured by rank 0 alone (so a scaling curve can be read off one run) ----
    single = None
    if n_gpus > 1 and gather in ("host", "ipc") and not weak:
        if rank == 0:
            ctx.set_interleave(1, 0)
            ctx.set_target_buffer(None, 0)
            k = max(3, min(steps, 10))
            s1 = frame_schedule(len(cams), k)
            ms1 = timed_frames(ctx, traversal, cams, s1, s1[:2])
            single = {"value": round(W * H * k / (ms1 / 1e3) / 1e6, 2), "unit": "Mrays/s", "steps": k,
                      "note": "same frame, same volume, one GPU (rank 0 alone, device-timed)"}
            ctx.set_interleave(n_gpus, rank)
        barrier()

    # ---- reduce over ranks: max time, sum rays ----
    region_all, e2e_all, rays_all = region_ms, e2e_s, my_rays
    kernel_ms_all = [float(np.mean(kernel_ms))]
    if n_gpus > 1:
        t = torch.tensor([region_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_all, e2e_all = t.tolist()
        r = torch.tensor([my_rays], dtype=torch.int64, device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays_all = int(r.item())
        km = torch.zeros(n_gpus, dtype=torch.float64, device="cuda")
        km[rank] = float(np.mean(kernel_ms))
        dist.all_reduce(km, op=dist.ReduceOp.SUM)
        kernel_ms_all = km.tolist()

    # ---- algorithmic bytes (instrumented pass, untimed) for the roofline of the dominant kernel ----
    if n_gpus > 1 and gather != "nccl":
        ctx.set_target_buffer(None, 0)
    alg_bytes, alg_steps, n_stat = algorithmic_bytes(ctx, traversal, cams, sched, my_rays)

    if rank != 0:
        if shared:
            dist.barrier()
            shared.close()
        if n_gpus > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = rays_all * steps / (region_all / 1e3) / 1e6
    e2e_value = rays_all * steps / e2e_all / 1e6
    mean_kernel_ms = float(np.mean(kernel_ms))
    kernel_name = kernel_name_for(xb, ctx, traversal)
    e2e = {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 36 * n_gpus,
           "d2h_bytes_per_step": e2e_d2h, "frames_per_s": round(steps / e2e_all, 2),
           "d2h_gb_per_s": round(e2e_d2h * steps / e2e_all / 1e9, 2)}
    e2e.update(e2e_info)
    result = {
        "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(region_all / steps, 5), "higher_is_better": True,
        "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": round(steps / (region_all / 1e3), 2),
        "config": {
            "workload": f"{workload}: {wl['desc']}", "traversal": traversal, "frame": f"{W}x{H}",
            "volume": f"{kind_name} {nx}x{ny}x{nz} seed {SEED}", "camera": wl["camera"], "emission": EMISSION,
            "frames": f"{steps} steps = script frames floor(i*{len(cams)}/{steps}) (whole path: exterior, fly-over, interior)",
            "grid_layout": grid_layout if traversal == "dda" else None,
            "partition": ("single region" if n_gpus == 1 else
                          f"16-row stripes round-robin over {n_gpus} GPUs, volume replicated; kernel-timed region: peer stores into "
                          f"rank 0's frame (CUDA IPC / NVLink); e2e: --gather {gather}"
                          if gather != "nccl" else f"{n_gpus} horizontal bands, NCCL send/recv gather to rank 0"),
            "l2": "inputs larger than L2 (volume resident in HBM exceeds 126 MB); camera changes every step",
        },
        "e2e": e2e,
        "gpu_launches": int(launches) * n_gpus,
        "clocks": clocks,
        "roofline": roofline_block(workload, traversal, kernel_name, alg_bytes, alg_steps, n_stat, mean_kernel_ms, n_gpus),
        "reference_stats": {"mray/s (rays / summed device ms, RenderStats.cpp:22-24)": round(
            rays_all / (sum(kernel_ms_all) * 1000.0), 2)},
    }
    if single:
        result["single_gpu_same_workload"] = single
    if n_gpus == 1 and extras:
        lb = compulsory_traffic(xb, ctx, traversal, cams, sched, kernel_ms)
        if lb:
            result["roofline"]["sector_lower_bound"] = lb

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample of the same workload ----
    if extras:
        try:
            if traversal == "dda":
                host_grid = ctx.download_grid().data  # the resident voxels themselves
            else:
                host_grid = np.zeros((1, 1, 1, 4), np.uint8)  # unused by the octree traversals
                if tree is None:
                    raise RuntimeError("node array of this volume is not brought back to the host")
            # about 10 s of CPU work: 12 frames spread over the script x 16 bands of 8 rows
            k = max(2, min(steps, 12))
            sub = [cam_tuple(cams, f) for f in frame_schedule(len(cams), k)]
            bands = oracle_sample_rows(H, bands=16, rows_per_band=8)
            mr, info = run_oracle_frames(traversal, (W, H), sub[:1] + sub, host_grid, tree, bands=bands)
            result["cpu_baseline"] = {"value": round(mr, 3), "unit": "Mrays/s", "cores": info["cores"],
                                      "kind": info["kind"], "sample": info["sample"],
                                      "seconds": round(info["seconds"], 2)}
            # parity spot check of the last e2e frame against the oracle (not timed)
            from oracle import xo
            cam = cam_tuple(cams, sched[-1])
            y0 = H // 2 - 8
            kw = dict(camera=cam, output=(0, y0, W, 16), display=display, emission=EMISSION, want_stats=False,
                      threads=host_threads())
            ref = (xo.render("dda", grid=host_grid, **kw) if traversal == "dda"
                   else xo.render(traversal, nodes=tree.nodes, side=tree.side, **kw))[0]
            d = np.abs(ref.astype(int) - last_frame[y0:y0 + 16].astype(int)).max(axis=-1)
            result["parity_check"] = {"rows": 16, "frame": int(sched[-1]), "within_1_of_255": float((d <= 1).mean()),
                                      "max_diff": int(d.max())}
        except Exception as e:  # the baseline must never take the GPU number down with it
            result["cpu_baseline"] = {"error": str(e)}

    if extras:
        try:
            result["per_config"] = per_config_block(xb, ctx, workload, traversal, result, steps, warmup, local_rank)
        except Exception as e:  # extras must never take the headline down with them
            result["per_config"] = {"error": str(e)}

    print(json.dumps(result), file=_JSON_OUT, flush=True)
    if shared:
        dist.barrier()
        shared.close()
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()


def per_config_block(xb, ctx, workload, traversal, headline, steps, warmup, device):
    """The other BASELINE configurations measured the same way (device-timed, whole camera path) in
    the same process; the headline's own entry is copied in.  Volumes are generated in HBM, octrees
    built on the GPU from the resident grid."""
    out = {}
    head = {"workload": headline["config"]["workload"], "traversal": traversal, "value": headline["value"],
            "unit": "Mrays/s", "ms_per_step": headline["ms_per_step"], "frames_per_s": headline["frames_per_s"],
            "steps": steps, "roofline": headline["roofline"]}
    out[workload if traversal == WORKLOADS[workload]["traversal"] else f"{workload}:{traversal}"] = head
    plan = [  # (volume key, [(config name, traversal)])
        (("tng", 2048, 2048, 2048), [("cfg4", "dda"), ("cfg4e", "esvo")]),
        (("tng", 1024, 1024, 1024), [("cfg3", "dda"), ("cfg3r", "svo-rope")]),
        (("bunny", 512, 361, 512), [("cfg1", "dda"), ("cfg2", "esvo"), ("cfg2", "svo-rope"), ("cfg2", "svo-df"),
                                    ("cfg2", "svo-naive"), ("cfg2", "dda")]),
    ]
    cur_volume = WORKLOADS[workload]["volume"]
    cur_tree = {"esvo": "sparse", "svo-df": "sparse", "svo-naive": "sparse", "svo-rope": "rope"}.get(traversal)
    for vol, entries in plan:
        todo = [(n, t) for n, t in entries
                if (n if t == WORKLOADS[n]["traversal"] else f"{n}:{t}") not in out]
        if not todo:
            continue
        if vol != cur_volume:
            kind_name, nx, ny, nz = vol
            ctx.synth_grid(xb.SYNTH_BUNNY if kind_name == "bunny" else xb.SYNTH_TNG, nx, ny, nz, SEED)
            cur_volume, cur_tree = vol, None
        dims = vol[1:]
        side = None
        # the rope tree serves every octree traversal (ropes live in leaf records the others never read)
        need_rope = any(t == "svo-rope" for _, t in todo)
        for name, t in sorted(todo, key=lambda e: e[1] != "dda"):
            if t != "dda":
                want = "rope" if need_rope else "sparse"
                if cur_tree != want and not (cur_tree == "rope" and want == "sparse"):
                    _, _, _, side = ctx.convert_resident_grid(
                        chan_diff=0, type=xb.TYPE_ROPE if want == "rope" else xb.TYPE_SPARSE, bind=True, want_nodes=False)
                    cur_tree = want
                elif side is None:
                    side = 1 << int(np.ceil(np.log2(max(dims))))
            cams = load_cameras(WORKLOADS[name]["camera"])
            key = name if t == WORKLOADS[name]["traversal"] else f"{name}:{t}"
            t0 = time.perf_counter()
            out[key] = measure_config(xb, ctx, name, t, WORKLOADS[name]["frame"], cams, steps, warmup, side, dims)
            log(f"[bench] per_config {key}: {out[key]['value']} Mrays/s ({time.perf_counter() - t0:.1f} s)")
    return out


def run_reference_arm(args, workload, wl, traversal, frame, cams, n_gpus, weak):
    """--impl reference: the reference's compute-shader text compiled as C++ (oracle/_ref, OpenMP) on
    the host cores, or the C restatement where that library is absent.  No GPU call on this arm: the
    volume comes from the host generator (bit-identical to the device generator)."""
    import xenodon_b200 as xb  # host-side formats / generators only
    kind_name, nx, ny, nz = wl["volume"]
    W, H = frame
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    threads = host_threads()
    t0 = time.perf_counter()
    log(f"[reference arm] generating {kind_name} {nx}x{ny}x{nz} on the host ({nx * ny * nz * 4 / 2**30:.1f} GiB) ...")
    grid = xb.Grid.synthetic(xb.SYNTH_BUNNY if kind_name == "bunny" else xb.SYNTH_TNG, nx, ny, nz, SEED)
    t_gen = time.perf_counter() - t0
    tree = None
    if traversal != "dda":
        log("[reference arm] convert --chan-diff 0 ...")
        tree, _ = xb.build_octree(grid, chan_diff=0, type=xb.TYPE_ROPE if traversal == "svo-rope" else xb.TYPE_SPARSE)
    # a step of this arm = a bounded row sample of one frame; the frames are the GPU arm's schedule
    # (thinned to at most 24 so the arm ends within minutes on the big volumes)
    sched = frame_schedule(len(cams), steps)
    n = max(1, min(steps, 24))
    pick = [sched[(i * len(sched)) // n] for i in range(n)]
    bands = oracle_sample_rows(H, bands=8, rows_per_band=16)
    cam_list = [cam_tuple(cams, f) for f in pick]
    warm = min(max(warmup, 0), 1)
    t1 = time.perf_counter()
    mr, info = run_oracle_frames(traversal, (W, H), cam_list[:warm] + cam_list, grid.data, tree, prefer_ref=True,
                                 warm=warm, bands=bands)
    wall = time.perf_counter() - t1
    rows = sum(r for _, r in bands)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": round(mr, 3), "unit": "Mrays/s", "n_gpus": n_gpus,
        "steps": steps, "warmup": warmup,
        "ms_per_step": round(info["seconds"] / max(info["frames"], 1) * 1e3 * (H / rows), 3),
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: {wl['desc']}", "traversal": traversal, "frame": f"{W}x{H}",
                   "volume": f"{kind_name} {nx}x{ny}x{nz} seed {SEED}", "camera": wl["camera"],
                   "emission": EMISSION,
                   "frames": f"{info['frames']} of the GPU arm's {steps} script frames floor(i*{len(cams)}/{steps})",
                   "note": "the reference's Vulkan build cannot run in this image (no loader/ICD/glslc); this arm "
                           "times its compute-shader text compiled as C++ (oracle/_ref, kind 'reference') or, where "
                           "that library is absent, the C restatement oracle/xn_oracle.c (kind 'port'); ms_per_step is "
                           "the sampled rows scaled to a whole frame"},
        "cpu_baseline": {"value": round(mr, 3), "unit": "Mrays/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "host_threads_available": threads,
                         "omp_num_threads_env_at_launch": os.environ.get("XN_LAUNCH_OMP", None)},
        "e2e": {"value": round(mr, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 1), "volume_generation_s": round(t_gen, 1),
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


if __name__ == "__main__":
    os.environ.setdefault("XN_LAUNCH_OMP", os.environ.get("OMP_NUM_THREADS", "unset"))
    main()
