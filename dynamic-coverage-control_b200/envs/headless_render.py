"""Headless stand-in for the reference's pyglet viewer (envs/mpe/multiagent/rendering.py, environment.py:209-330).

The reference draws, per env, on a 700x700 window with camera range [-2, 2]^2: PoI discs (grey -> green with
energy, CoverageWorld.py:172-173), UAV discs, a translucent cover disc (r_cover) and comm disc (r_comm) per UAV
(scenarios/coverage.py:65-68 colours) and a line per communicating UAV pair.  This module rasterises the same
scene from the compact state (positions, PoI energy, adjacency bitmasks) with NumPy — no display, no pyglet —
and the trajectory recorder keeps the raw state so that plots (connectivity rate, coverage curves) can be made
offline.  Host-side visualisation only: not on the hot path.
"""
import numpy as np

CAM_RANGE = 2.0
AGENT_COLOR = np.array([0.05, 0.15, 0.05])
COVER_COLOR = np.array([0.05, 0.25, 0.05])
COMM_COLOR = np.array([0.05, 0.35, 0.05])
ENTITY_SIZE = 0.02


def _disc(img, yy, xx, cx, cy, r, color, alpha):
    m = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
    img[m] = img[m] * (1.0 - alpha) + color * alpha


def _line(img, size, p, q, color):
    n = int(max(abs(q[0] - p[0]), abs(q[1] - p[1]))) + 1
    xs = np.clip(np.round(np.linspace(p[0], q[0], n)).astype(int), 0, size - 1)
    ys = np.clip(np.round(np.linspace(p[1], q[1], n)).astype(int), 0, size - 1)
    img[ys, xs] = color


def rasterize(pos, poi_xy, energy, adj_rows=None, r_cover=0.2, r_comm=0.4, m_energy=5.0, size=350):
    """One RGB frame (size, size, 3) uint8.  pos (N,2), poi_xy (M,2), energy (M,), adj_rows (N,) uint32 bitmasks."""
    pos = np.asarray(pos, dtype=np.float64)
    poi_xy = np.asarray(poi_xy, dtype=np.float64)
    energy = np.asarray(energy, dtype=np.float64)
    img = np.ones((size, size, 3), dtype=np.float64)
    scale = size / (2.0 * CAM_RANGE)
    yy, xx = np.mgrid[0:size, 0:size]

    def px(p):   # world -> pixel (y up)
        return (p[0] + CAM_RANGE) * scale, (CAM_RANGE - p[1]) * scale

    for i in range(pos.shape[0]):
        cx, cy = px(pos[i])
        _disc(img, yy, xx, cx, cy, r_comm * scale, COMM_COLOR, 0.15)
    for i in range(pos.shape[0]):
        cx, cy = px(pos[i])
        _disc(img, yy, xx, cx, cy, r_cover * scale, COVER_COLOR, 0.15)
    for j in range(poi_xy.shape[0]):
        done = energy[j] >= m_energy
        g = 1.0 if done else 0.25 + min(energy[j] / m_energy, 1.0) * 0.75
        cx, cy = px(poi_xy[j])
        _disc(img, yy, xx, cx, cy, max(ENTITY_SIZE * scale, 1.5), np.array([0.25, g, 0.25]), 1.0)
    if adj_rows is not None:
        rows = np.asarray(adj_rows).astype(np.uint32)
        for a in range(pos.shape[0]):
            for b in range(a + 1, pos.shape[0]):
                if (int(rows[a]) >> b) & 1:
                    _line(img, size, px(pos[a]), px(pos[b]), np.array([0.2, 0.2, 0.8]))
    for i in range(pos.shape[0]):
        cx, cy = px(pos[i])
        _disc(img, yy, xx, cx, cy, max(ENTITY_SIZE * scale, 2.0), AGENT_COLOR, 0.5)
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


class TrajectoryRecorder:
    """Per-step compact state of a few env instances: positions, velocities, PoI energy, connect bits, adjacency,
    coverage rate, reward.  `save(path)` writes one .npz; `frames()` rasterises env 0."""

    def __init__(self, poi_xy, r_cover, r_comm):
        self.poi_xy, self.r_cover, self.r_comm = np.asarray(poi_xy), float(r_cover), float(r_comm)
        self.steps = []

    def add(self, pos_vel, energy, connect_bits, adj_rows, coverage_rate, reward):
        self.steps.append(dict(pos_vel=np.array(pos_vel), energy=np.array(energy), connect_bits=np.array(connect_bits),
                               adj=None if adj_rows is None else np.array(adj_rows),
                               coverage_rate=np.array(coverage_rate), reward=np.array(reward)))

    def arrays(self):
        out = {k: np.stack([s[k] for s in self.steps]) for k in ("pos_vel", "energy", "connect_bits", "coverage_rate", "reward")}
        if self.steps and self.steps[0]["adj"] is not None:
            out["adj"] = np.stack([s["adj"] for s in self.steps])
        out["poi_xy"] = self.poi_xy
        out["r_cover"], out["r_comm"] = np.array(self.r_cover), np.array(self.r_comm)
        return out

    def connectivity_rate(self):
        cb = np.stack([s["connect_bits"] for s in self.steps])
        return float((cb & 1).mean())

    def frames(self, env=0, size=350):
        return [rasterize(s["pos_vel"][env, :, :2], self.poi_xy, s["energy"][env],
                          None if s["adj"] is None else s["adj"][env], self.r_cover, self.r_comm, size=size)
                for s in self.steps]

    def save(self, path):
        np.savez_compressed(path, **self.arrays())

    def save_gif(self, path, env=0, size=350, duration_ms=100):
        from PIL import Image     # optional: only for the GIF the reference writes with imageio (learner.py:204-210)
        ims = [Image.fromarray(f) for f in self.frames(env, size)]
        if ims:
            ims[0].save(path, save_all=True, append_images=ims[1:], duration=duration_ms, loop=0)
