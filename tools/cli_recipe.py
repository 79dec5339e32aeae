#!/usr/bin/env python3
"""The reference's own benchmark recipe (README.md:85-89 of the reference) through the drop-in CLI:
    xenodon render --headless headless.conf <volume> --camera camera.txt -e 10 --discard-output --stats-output <file>
on the bunny-shaped 512x361x512 volume (cfg1 / cfg2), as a TIFF stack for the DDA and -- after
`xenodon convert [--rope]` on the GPU -- as .svo files for the four octree traversals.  Keeps the stats
files' summaries (the reference's RenderStatsAccumulator::save format).   usage: cli_recipe.py <out prefix>"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xenodon_b200 as xb  # noqa: E402
from xenodon_b200 import cameras  # noqa: E402


def run(*args):
    t0 = time.perf_counter()
    r = subprocess.run([xb.CLI_PATH, *map(str, args)], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stdout + r.stderr)
    return r.stdout + r.stderr, time.perf_counter() - t0


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cli"
    d = tempfile.mkdtemp(prefix="xn_cli_")
    tif, conf, cam = os.path.join(d, "bunny.tif"), os.path.join(d, "headless.conf"), os.path.join(d, "camera.txt")
    xb.Grid.synthetic(xb.SYNTH_BUNNY, 512, 361, 512, 1729).save_tiff(tif)
    open(conf, "w").write("device {\n    vkindex = 0\n    offset = (0, 0)\n    extent = (1920, 1080)\n}\n")
    open(cam, "w").write(cameras.to_text(cameras.camera_benchmark()))
    report = []
    out, s = run("convert", tif, os.path.join(d, "bunny.svo"))
    report.append(f"# xenodon convert bunny.tif bunny.svo: {s:.2f} s wall\n" + "".join(
        line + "\n" for line in out.splitlines() if "nodes" in line or "Built" in line))
    out, s = run("convert", "--rope", tif, os.path.join(d, "bunny-rope.svo"))
    report.append(f"# xenodon convert --rope bunny.tif bunny-rope.svo: {s:.2f} s wall\n")
    for shader, vol in (("dda", tif), ("esvo", "bunny.svo"), ("svo-naive", "bunny.svo"), ("svo-df", "bunny.svo"),
                        ("svo-rope", "bunny-rope.svo")):
        stats = os.path.join(d, f"stats-{shader}.txt")
        _, s = run("render", "--headless", conf, os.path.join(d, vol) if vol != tif else tif, "--camera", cam, "-e", "10",
                   "-s", shader, "--discard-output", "--stats-output", stats, "-q")
        lines = open(stats).read().splitlines()
        head = [ln for ln in lines if not ln.startswith("frame ")]
        report.append(f"# xenodon render --headless headless.conf {os.path.basename(vol)} --camera camera.txt -e 10 -s {shader} "
                      f"--discard-output --stats-output stats.txt   ({s:.2f} s wall incl. loading)\n" + "\n".join(head) + "\n")
    open(prefix + "_stats.txt", "w").write("\n".join(report))
    print("\n".join(report))


if __name__ == "__main__":
    main()
