"""Seeded synthetic workloads of SURVEY.md §8(d) (the configs named in BASELINE.json), as ``ModelSpec`` objects.

Common: voxel 0.01 m; material A E=1e6 Pa rho=1e3 cte=+0.01, B E=5e6 Pa rho=1.5e3 cte=-0.01, passive C E=1e6 cte=0;
nu=0; uStatic=1, uDynamic=0.8; BondDampingZ=1, ColDampingZ=0.8, SlowDampingZ=0.01; gravity -9.81, floor on;
TempAmplitude=20, TempPeriod=0.2, VaryTemp on; DtFrac=0.9; PRNG = splitmix64.
"""
import numpy as np

from . import abi
from .model import ModelSpec


def splitmix64(seed):
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    return nxt


def _u01(r):
    return (r() >> 11) / float(1 << 53)


def add_abc_materials(spec, sticky=False):
    spec.add_material(name="A", elastic_mod=1e6, density=1e3, cte=0.01, u_static=1.0, u_dynamic=0.8, red=1.0, green=0.0, blue=0.0,
                      sticky=int(sticky))
    spec.add_material(name="B", elastic_mod=5e6, density=1.5e3, cte=-0.01, u_static=1.0, u_dynamic=0.8, red=0.0, green=1.0, blue=0.0)
    spec.add_material(name="C", elastic_mod=1e6, density=1e3, cte=0.0, u_static=1.0, u_dynamic=0.8, red=0.0, green=0.0, blue=1.0)


def _common(spec, collisions=0):
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01, temp_enabled=1, vary_temp_enabled=1,
                 temp_amplitude=20.0, temp_period=0.2)
    spec.set_options(enable_collision=collisions)


def body_spec(n=(20, 20, 20), seed=42, name="c2", fill=1.0, keep_largest=False, min_voxels=0):
    """A/B/C body on an nx*ny*nz lattice: material ~ U{A,B,C} (seed), phase ~ U[0,1) (seed+1), cell kept with
    probability `fill` (seed+2); optionally only the largest 6-connected component is kept."""
    nx, ny, nz = n
    spec = ModelSpec(0.01, name)
    add_abc_materials(spec)
    _common(spec)
    r, r2, r3 = splitmix64(seed), splitmix64(seed + 1), splitmix64(seed + 2)
    st = np.zeros((nz, ny, nx), np.uint8)
    ph = np.zeros((nz, ny, nx))
    for z in range(nz):
        for y in range(ny):
            for x in range(nx):
                m = 1 + r() % 3
                ph[z, y, x] = _u01(r2)
                if _u01(r3) < fill:
                    st[z, y, x] = m
    if keep_largest:
        st = _largest_component(st)
        if (st > 0).sum() < min_voxels:  # degenerate draw: fall back to the full lattice of the same materials
            return body_spec(n, seed, name, 1.0, False, 0)
        # drop the body onto the floor: shift down so the lowest filled layer is z=0
        zs = np.nonzero(st.any(axis=(1, 2)))[0]
        if zs[0] > 0:
            st = np.concatenate([st[zs[0]:], np.zeros((zs[0],) + st.shape[1:], np.uint8)])
            ph = np.concatenate([ph[zs[0]:], np.zeros((zs[0],) + ph.shape[1:])])
    spec.set_structure(st, phase_offset=ph)
    return spec


def _largest_component(st):
    filled = st > 0
    lab = -np.ones(st.shape, np.int64)
    best, best_n, cur = -1, 0, 0
    nz, ny, nx = st.shape
    for z0, y0, x0 in zip(*np.nonzero(filled)):
        if lab[z0, y0, x0] >= 0:
            continue
        stack, cnt = [(z0, y0, x0)], 0
        lab[z0, y0, x0] = cur
        while stack:
            z, y, x = stack.pop()
            cnt += 1
            for dz, dy, dx in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
                a, b, c = z + dz, y + dy, x + dx
                if 0 <= a < nz and 0 <= b < ny and 0 <= c < nx and filled[a, b, c] and lab[a, b, c] < 0:
                    lab[a, b, c] = cur
                    stack.append((a, b, c))
        if cnt > best_n:
            best, best_n = cur, cnt
        cur += 1
    out = st.copy()
    out[lab != best] = 0
    return out


def c2_spec():
    """Config 2: single 20x20x20 multi-material actuated body, collisions off (100,000 steps in the full run)."""
    return body_spec((20, 20, 20), seed=42, name="c2_20x20x20")


def c3_spec(k):
    """Config 3, robot k: 10^3 lattice filled with p=0.7, largest component (>=100 voxels), stop t>1 s, fitness sqrt(x^2+y^2)."""
    spec = body_spec((10, 10, 10), seed=1000 + k, name="c3_robot_%04d" % k, fill=0.7, keep_largest=True, min_voxels=100)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 1.0)))
    spec.set_program(abi.PROG_FITNESS, ("SQRT", ("ADD", ("MUL", ("VAR", "x"), ("VAR", "x")), ("MUL", ("VAR", "y"), ("VAR", "y")))))
    return spec


def c4_spec(grid=(8, 8, 8), body=4, name="c4_pile"):
    """Config 4: grid of body^3 sticky bodies (one sticky material with a failure stress), 1 empty cell between
    bodies horizontally and 2 vertically, dropped onto the floor; collisions + attach + detach."""
    gx, gy, gz = grid
    spec = ModelSpec(0.01, name)
    spec.add_material(name="S", mat_model=1, elastic_mod=1e6, fail_stress=3.5e5, density=1e3, cte=0.01, u_static=1.0, u_dynamic=0.8,
                      sticky=1, red=1.0, green=0.6, blue=0.1)
    _common(spec, collisions=1)
    spec.set_options(enable_collision=1, enable_attach=1, enable_detach=1)
    px, pz = body + 1, body + 2
    st = np.zeros((gz * pz, gy * px, gx * px), np.uint8)
    ph = np.zeros(st.shape)
    r = splitmix64(77)
    for k in range(gz):
        for j in range(gy):
            for i in range(gx):
                st[k * pz:k * pz + body, j * px:j * px + body, i * px:i * px + body] = 1
                ph[k * pz:k * pz + body, j * px:j * px + body, i * px:i * px + body] = _u01(r)
    spec.set_structure(st, phase_offset=ph)
    return spec


def c5_spec(n=(200, 200, 100)):
    """Config 5: single-material body with a checkerboard phase offset {0, 0.5}."""
    nx, ny, nz = n
    spec = ModelSpec(0.01, "c5_%dx%dx%d" % n)
    add_abc_materials(spec)
    _common(spec)
    st = np.ones((nz, ny, nx), np.uint8)
    z, y, x = np.indices(st.shape)
    ph = 0.5 * ((x + y + z) % 2)
    spec.set_structure(st, phase_offset=ph.astype(np.float64))
    return spec


def alg_bytes_per_voxel_step(n_voxels, n_links):
    """Algorithmic HBM bytes per voxel-step: 228 B per voxel + 184 B per link."""
    return 228.0 + 184.0 * (n_links / float(n_voxels))
