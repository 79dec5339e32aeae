import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import xenodon_b200 as xb
from xenodon_b200 import cameras
cams = cameras.camera_benchmark()
w, h = 1280, 720
def cam(f): return (tuple(f[0]), tuple(f[1]), tuple(f[2]))
for n in (256, 512, 1024):
    for kind in (xb.SYNTH_TNG, xb.SYNTH_BUNNY):
        ctx = xb.Context(0)
        ctx.set_grid_layout(xb.LAYOUT_LINEAR)
        ctx.synth_grid(kind, n, n, n)
        ctx.set_target((0, 0, w, h)); ctx.set_precision(True)
        ctx.set_params((1, 1, 1), (n, n, n), 1.0)
        c = cam(cams[10])
        ctx.render("dda", c); ctx.sync(); dda = ctx.download().astype(int)
        _, st, count, side = ctx.convert_resident_grid(chan_diff=0, type=xb.TYPE_ROPE, bind=True)
        ctx.set_params((1, 1, 1), (side,) * 3, 1.0)
        res = {}
        for t in ("esvo", "svo-rope", "svo-naive", "svo-df"):
            ctx.render(t, c); ctx.sync(); res[t] = ctx.download().astype(int)
        line = [f"n={n} kind={kind} nodes={count}"]
        for t, img in res.items():
            d = np.abs(img - dda).max(axis=-1)
            line.append(f"{t}:dda within1={float((d<=1).mean()):.4f} max={int(d.max())}")
        d = np.abs(res['svo-rope'] - res['esvo']).max(axis=-1)
        line.append(f"rope:esvo within1={float((d<=1).mean()):.4f} max={int(d.max())}")
        print(" | ".join(line), flush=True)
        ctx.close()
