"""Seeded synthetic inputs for the RBF deformation path.

The reference ships no sample scenes, so every test and bench line uses these generators:
a jittered height-field "face" mesh with quad topology (what input 0 of the SOP would carry), a control rig
of N markers sampled on it (inputs 1 and 2, SOP_FaceDeform.cpp:228-229) and per-frame marker displacements.
All arrays are float32 like Houdini's P attribute; all randomness comes from numpy.random.default_rng(seed).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SEED_MESH, SEED_CTRL, SEED_DELTA = 1, 2, 3


def _height(u, v):
    return 0.3 * np.exp(-4.0 * (u * u + v * v))


def _normal(u, v):
    # gradient of the height field z = 0.3 exp(-4(u^2+v^2))
    z = _height(u, v)
    n = np.stack([8.0 * u * z, 8.0 * v * z, np.ones_like(u)], axis=-1)
    return n / np.linalg.norm(n, axis=-1, keepdims=True)


@dataclass
class Mesh:
    P: np.ndarray          # (V, 3) float32
    poly_off: np.ndarray   # (npoly + 1,) int32 CSR offsets
    poly_vtx: np.ndarray   # (4 * npoly,) int32 quad corners
    tangentu: np.ndarray   # (V, 3) float32 (not normalised, like PolyFrame output may be)
    tangentv: np.ndarray
    N: np.ndarray
    bbox_diag: float


def grid_dims(V: int):
    nu = int(math.ceil(math.sqrt(V)))
    nv = int(math.ceil(V / nu))
    return nu, nv


def face_mesh(V: int, seed: int = SEED_MESH, topology: bool = True) -> Mesh:
    """V points of a nu x nv jittered grid over [-1, 1]^2 lifted onto the height field; first V points kept."""
    rng = np.random.default_rng(seed)
    nu, nv = grid_dims(V)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    iu = iu.reshape(-1)[:V]
    iv = iv.reshape(-1)[:V]
    du, dv = 2.0 / max(nu - 1, 1), 2.0 / max(nv - 1, 1)
    u = -1.0 + iu * du + rng.uniform(-0.25, 0.25, V) * du
    v = -1.0 + iv * dv + rng.uniform(-0.25, 0.25, V) * dv
    z = _height(u, v) + 0.02 * rng.standard_normal(V) * min(du, dv) * 10.0
    P = np.stack([u, v, z], axis=1).astype(np.float32)
    if topology:
        a = (np.arange(nu - 1)[:, None] * nv + np.arange(nv - 1)[None, :]).reshape(-1)
        quads = np.stack([a, a + nv, a + nv + 1, a + 1], axis=1)
        quads = quads[(quads < V).all(axis=1)]
        poly_vtx = quads.reshape(-1).astype(np.int32)
        poly_off = (4 * np.arange(quads.shape[0] + 1)).astype(np.int32)
    else:
        poly_vtx = np.zeros(0, np.int32)
        poly_off = np.zeros(1, np.int32)
    n = _normal(u, v)
    zc = _height(u, v)
    tu = np.stack([np.ones(V), np.zeros(V), -8.0 * u * zc], axis=1)
    tv = np.stack([np.zeros(V), np.ones(V), -8.0 * v * zc], axis=1)
    scale = rng.uniform(0.5, 2.0, (V, 1))
    lo, hi = P.min(axis=0).astype(np.float64), P.max(axis=0).astype(np.float64)
    return Mesh(P, poly_off, poly_vtx, (tu * scale).astype(np.float32), (tv * scale).astype(np.float32),
                (n * scale).astype(np.float32), float(np.linalg.norm(hi - lo)))


@dataclass
class Rig:
    rest: np.ndarray       # (N, 3) float32
    normals: np.ndarray    # (N, 3) float64 (surface normals at the markers)
    prim_off: np.ndarray   # CSR triangles over the markers (Delaunay in the parameter plane)
    prim_vtx: np.ndarray
    spacing: float         # mean nearest-neighbour distance


def control_rig(N: int, seed: int = SEED_CTRL, prims: bool = False) -> Rig:
    """N never-coincident markers: one per cell of a coarse m x m grid (m = ceil(sqrt N)), jittered by 0.25 cell."""
    rng = np.random.default_rng(seed)
    m = int(math.ceil(math.sqrt(N)))
    cells = rng.permutation(m * m)[:N]
    cells.sort()
    cu, cv = cells // m, cells % m
    h = 1.9 / m
    u = -0.95 + (cu + 0.5) * h + rng.uniform(-0.25, 0.25, N) * h
    v = -0.95 + (cv + 0.5) * h + rng.uniform(-0.25, 0.25, N) * h
    z = _height(u, v) + 0.002
    rest = np.stack([u, v, z], axis=1).astype(np.float32)
    r64 = rest.astype(np.float64)
    if N > 1:
        from scipy.spatial import cKDTree
        d, _ = cKDTree(r64).query(r64, k=2)
        spacing = float(d[:, 1].mean())
    else:
        spacing = 1.0
    if prims and N >= 3:
        from scipy.spatial import Delaunay
        tri = Delaunay(np.stack([u, v], axis=1)).simplices.astype(np.int32)
        prim_vtx = tri.reshape(-1)
        prim_off = (3 * np.arange(tri.shape[0] + 1)).astype(np.int32)
    else:
        prim_vtx = np.zeros(0, np.int32)
        prim_off = np.zeros(1, np.int32)
    return Rig(rest, _normal(u, v), prim_off, prim_vtx, spacing)


def deformed_rig(rig: Rig, F: int, seed: int = SEED_DELTA) -> np.ndarray:
    """(F, N, 3) float32 deformed marker positions: rest + 0.05 sin(2 pi f / F + phase_i) n_i + 0.01 N(0,1)."""
    rng = np.random.default_rng(seed)
    N = rig.rest.shape[0]
    phase = rng.uniform(0.0, 2.0 * np.pi, N)
    out = np.empty((F, N, 3), np.float32)
    for f in range(F):
        amp = 0.05 * np.sin(2.0 * np.pi * f / F + phase)
        delta = amp[:, None] * rig.normals + 0.01 * rng.standard_normal((N, 3))
        out[f] = rig.rest + delta.astype(np.float32)
    return out


# BASELINE.json configs -> concrete shapes ("Config shorthand")
CONFIGS = {
    "C1": dict(N=64, V=10_000, F=1, kernel="gaussian", term="linear"),
    "C2": dict(N=256, V=100_000, F=240, kernel="gaussian", term="linear"),
    "C3": dict(N=2048, V=1_000_000, F=1, kernel="multiquadric", term="linear"),
    "C3t": dict(N=2048, V=1_000_000, F=1, kernel="thin_plate", term="linear"),
    "C4": dict(N=8192, V=0, F=1, kernel="gaussian", term="linear"),
    "C5": dict(N=4096, V=16_000_000, F=1000, kernel="gaussian", term="linear"),
}

KERNELS = {"gaussian": 0, "multiquadric": 1, "thin_plate": 2}
TERMS = {"linear": 0, "const": 1, "zero": 2}


def default_radius(kernel: str, spacing: float) -> float:
    """Gaussian R = 2 x mean NN spacing, multiquadric R = mean spacing, thin plate has no parameter."""
    if kernel == "gaussian":
        return 2.0 * spacing
    if kernel == "multiquadric":
        return spacing
    return 1.0
