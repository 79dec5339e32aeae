"""Camera scripts of the benchmark configurations, generated from their closed forms.

The reference ships three camera files (camera-single.txt, camera.txt, camera-rotate.txt;
format: reference src/camera/ScriptCameraController.cpp:17-41).  They are data, not
code, and are not copied into this repository; the paths are regenerated here from the
formulas they follow (tests/test_host_formats.py checks them against the reference's
files where the reference checkout is present):

  camera-single : one frame looking down +z from (0.5, 0.5, -1.5).
  camera-rotate : 150 frames, yaw orbit of one full turn at distance 2 from the centre.
  camera        : 150 frames = 50 of a half-turn yaw orbit at distance 2, 50 pitching
                  over the top by pi while zooming from distance 2 to 0.5, 50 of a full
                  turn at distance 0.5 (origin inside the volume, upside-down).
Values are rounded to 6 significant digits like the files.
"""
from __future__ import annotations

import math

import numpy as np

CENTRE = (0.5, 0.5, 0.5)


def _frame(fwd, up, dist):
    pos = tuple(c - dist * f for c, f in zip(CENTRE, fwd))
    return [float("%g" % v) for v in (*fwd, *up, *pos)]


def camera_single() -> np.ndarray:
    return np.asarray([_frame((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 2.0)], dtype=np.float32).reshape(-1, 3, 3)


def camera_rotate(frames: int = 150) -> np.ndarray:
    out = []
    for i in range(frames):
        t = 2 * math.pi * i / frames
        out.append(_frame((math.sin(t), 0.0, math.cos(t)), (0.0, 1.0, 0.0), 2.0))
    return np.asarray(out, dtype=np.float32).reshape(-1, 3, 3)


def camera_benchmark() -> np.ndarray:
    out = []
    step = 2 * math.pi / 100
    for i in range(50):  # half-turn yaw orbit
        t = step * i
        out.append(_frame((math.sin(t), 0.0, math.cos(t)), (0.0, 1.0, 0.0), 2.0))
    t = step * 49
    for j in range(1, 51):  # over the top, zooming in
        p = step * j
        fwd = (math.sin(t) * math.cos(p), -math.sin(p), math.cos(t) * math.cos(p))
        up = (math.sin(t) * math.sin(p), math.cos(p), math.cos(t) * math.sin(p))
        out.append(_frame(fwd, up, 2.0 - 0.03 * j))
    for k in range(1, 51):  # full turn inside the volume
        a = -step - 2 * step * k
        out.append(_frame((math.sin(a), 0.0, math.cos(a)), (0.0, -1.0, 0.0), 0.5))
    return np.asarray(out, dtype=np.float32).reshape(-1, 3, 3)


def to_text(frames: np.ndarray) -> str:
    return "\n".join(" ".join("%g" % v for v in f.reshape(-1)) for f in frames)


SCRIPTS = {"camera-single": camera_single, "camera-rotate": camera_rotate, "camera": camera_benchmark}
