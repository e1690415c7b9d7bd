"""Seeded synthetic inputs for the five BASELINE.json configs.

The reference ships no datasets (its `dataset/` is git-ignored), so every test,
bench and profile in this repo runs on inputs generated here.  Everything is a
pure function of (config, sequence seed, frame number): seed = cfg*1000 + frame
(+ 100000*sequence for config 5), numpy PCG64.

  cfg 1  8x8 L5 depth matrices (int mm) + IMU params, JSON wire format of
         src/main.c:44-63,161-177
  cfg 2  16x1800 VLP-16-shaped clouds with *integer* mm xyz, CSV wire format of
         src/main.c:86-99
  cfg 3  64x2048 OS1-64-shaped clouds, fp64 xyz
  cfg 4  1 M-point map + 131 072 queries (uniform / clustered variants)
  cfg 5  8 independent cfg-3 sequences; 10 M-point map
"""
from __future__ import annotations

import json

import numpy as np

ROOM_HALF_X = 10000.0  # 20 m x 12 m box room, mm
ROOM_HALF_Y = 6000.0
ROOM_FLOOR = -1500.0
ROOM_CEIL = 2000.0


def _ray_dirs(rows: int, cols: int, elev_lo_deg: float, elev_hi_deg: float) -> np.ndarray:
    az = 2.0 * np.pi * np.arange(cols) / cols
    el = np.deg2rad(np.linspace(elev_lo_deg, elev_hi_deg, rows)) if rows > 1 else np.zeros(1)
    ce = np.cos(el)[:, None]
    d = np.empty((rows, cols, 3), dtype=np.float64)
    d[..., 0] = ce * np.cos(az)[None, :]
    d[..., 1] = ce * np.sin(az)[None, :]
    d[..., 2] = np.sin(el)[:, None] * np.ones((1, cols))
    return d


def room_frame(rows: int, cols: int, frame: int, *, cfg: int = 3, seq: int = 0,
               elev=(-16.0, 16.0), integer_mm: bool = False, invalid_frac: float = 0.0,
               dirs: np.ndarray | None = None) -> np.ndarray:
    """One lidar-frame range image [rows, cols, 3] (mm, fp64) of the box room with pillars
    every 64 columns (x0.6 range), +-10 mm uniform range noise; the sensor advances
    50 mm/frame along +x.  invalid_frac > 0 zeroes that fraction of returns to (0,0,0),
    which is what the reference's ingest produces for invalid data (pointcloud.c:24-27)."""
    rng = np.random.default_rng(cfg * 1000 + frame + 100000 * seq)
    if dirs is None:
        dirs = _ray_dirs(rows, cols, *elev)
    sx = -4000.0 + 50.0 * frame  # sensor position inside the room
    sy = 500.0 * seq
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = np.where(dirs[..., 0] > 0, (ROOM_HALF_X - sx) / dirs[..., 0], (-ROOM_HALF_X - sx) / dirs[..., 0])
        ty = np.where(dirs[..., 1] > 0, (ROOM_HALF_Y - sy) / dirs[..., 1], (-ROOM_HALF_Y - sy) / dirs[..., 1])
        tz = np.where(dirs[..., 2] > 0, ROOM_CEIL / dirs[..., 2], ROOM_FLOOR / dirs[..., 2])
    tx = np.where(np.isfinite(tx) & (tx > 0), tx, np.inf)
    ty = np.where(np.isfinite(ty) & (ty > 0), ty, np.inf)
    tz = np.where(np.isfinite(tz) & (tz > 0), tz, np.inf)
    rng_mm = np.minimum(np.minimum(tx, ty), tz)
    pillar = (np.arange(cols) % 64) < 3
    rng_mm = np.where(pillar[None, :], rng_mm * 0.6, rng_mm)
    rng_mm = rng_mm + rng.uniform(-10.0, 10.0, size=rng_mm.shape)
    pts = dirs * rng_mm[..., None]
    if integer_mm:
        pts = np.rint(pts)
    if invalid_frac > 0.0:
        bad = rng.random(size=rng_mm.shape) < invalid_frac
        pts[bad] = 0.0
    return np.ascontiguousarray(pts, dtype=np.float64)


def room_sequence(rows: int, cols: int, n_frames: int, *, cfg: int = 3, seq: int = 0,
                  elev=(-16.0, 16.0), integer_mm: bool = False, start: int = 0) -> np.ndarray:
    dirs = _ray_dirs(rows, cols, *elev)
    out = np.empty((n_frames, rows, cols, 3), dtype=np.float64)
    for f in range(n_frames):
        out[f] = room_frame(rows, cols, start + f, cfg=cfg, seq=seq, elev=elev,
                            integer_mm=integer_mm, dirs=dirs)
    return out


def true_pose(frame: int, seq: int = 0) -> np.ndarray:
    """Ground-truth sensor pose [x,y,z,roll,pitch,yaw] (mm, deg) relative to frame 0."""
    return np.array([50.0 * frame, 0.0, 0.0, 0.0, 0.0, 0.0], dtype=np.float64)


# ---------------------------------------------------------------- config 1 --
def l5_depth_frame(frame: int, rows: int = 8, cols: int = 8) -> np.ndarray:
    """Wall at 1500 mm with a 3x4-pixel box at 800 mm, approaching 20 mm/frame, +-3 mm noise."""
    rng = np.random.default_rng(1 * 1000 + frame)
    d = np.full((rows, cols), 1500 - 20 * frame, dtype=np.int64)
    r0, c0 = min(2, rows - 1), min(3, cols - 1)
    d[r0:r0 + 3, c0:c0 + 4] = 800 - 20 * frame
    d = d + rng.integers(-3, 4, size=d.shape)
    return d.astype(np.int32)


def l5_json(n_frames: int, rows: int = 8, cols: int = 8) -> str:
    """`parsed_data.json` of config 1.  params are written with a decimal point because
    json_real_value() returns 0 for JSON integers (src/main.c:171-176)."""
    frames = []
    for f in range(n_frames):
        d = l5_depth_frame(f, rows, cols)
        frames.append({
            "time_main": 1000 + 100 * f,
            "distance": [int(v) for v in d.reshape(-1)],
            "params": [0.0, 0.0, 0.0, round(0.02 * f, 6), 0.0, 0.0],
        })
    txt = json.dumps(frames)
    return txt


# ---------------------------------------------------------------- config 2 --
def l9_sequence(n_frames: int = 10, rows: int = 16, cols: int = 1800) -> np.ndarray:
    return room_sequence(rows, cols, n_frames, cfg=2, elev=(-15.0, 15.0), integer_mm=True)


def l9_csv(frames: np.ndarray) -> str:
    """`parsed_data.csv`: header + frame,row,col,x,y,z,conf (src/main.c:86-99)."""
    n, rows, cols, _ = frames.shape
    lines = ["frame,row,col,x,y,z,conf"]
    for f in range(n):
        for r in range(rows):
            for c in range(cols):
                x, y, z = frames[f, r, c]
                lines.append(f"{f},{r},{c},{int(x)},{int(y)},{int(z)},100")
    return "\n".join(lines) + "\n"


# ------------------------------------------------------------- config 4/5 --
def map_points(n: int, *, variant: str = "uniform", seed: int = 4000) -> np.ndarray:
    """n map points in a 100 m x 100 m x 10 m volume (mm).  'clustered' puts them on the
    surfaces of 64 boxes (room-like), which gives many equal coordinates on one axis."""
    rng = np.random.default_rng(seed)
    if variant == "uniform":
        p = rng.random((n, 3))
        p *= np.array([100000.0, 100000.0, 10000.0])
        return p
    if variant == "clustered":
        centers = rng.random((64, 3)) * np.array([100000.0, 100000.0, 10000.0])
        half = rng.uniform(500.0, 4000.0, size=(64, 3))
        which = rng.integers(0, 64, size=n)
        face = rng.integers(0, 6, size=n)
        u = rng.uniform(-1.0, 1.0, size=(n, 3))
        axis = face // 2
        sign = np.where(face % 2 == 0, -1.0, 1.0)
        u[np.arange(n), axis] = sign
        p = centers[which] + half[which] * u
        p += rng.normal(0.0, 5.0, size=p.shape)
        return p
    raise ValueError(variant)


def accumulated_map(n_frames: int = 8, rows: int = 64, cols: int = 2048, *, start: int = 0):
    """BASELINE config 4 as a SLAM run produces it: the global clouds of n_frames consecutive frames
    (each moved by its true pose) concatenated in mapping order -- 8 x 64 x 2048 = 1 048 576 points --
    and the next frame, moved by its predicted pose, as the 131 072 queries in image order."""
    frames = room_sequence(rows, cols, n_frames + 1, start=start)
    parts = [frames[f].reshape(-1, 3) + true_pose(start + f)[:3] for f in range(n_frames)]
    queries = frames[n_frames].reshape(-1, 3) + true_pose(start + n_frames)[:3]
    return np.ascontiguousarray(np.concatenate(parts)), np.ascontiguousarray(queries)


def map_queries(points: np.ndarray, nq: int, *, variant: str = "jitter", seed: int = 4001) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if variant == "jitter":
        pick = rng.integers(0, points.shape[0], size=nq)
        return points[pick] + rng.normal(0.0, 50.0, size=(nq, 3))
    if variant == "uniform":
        lo, hi = points.min(axis=0), points.max(axis=0)
        return lo + rng.random((nq, 3)) * (hi - lo)
    raise ValueError(variant)
