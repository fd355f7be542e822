"""ctypes front end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (accelerated-3d-acoustic-fdtd-kernel_b200) never does.

Two back ends with the same call surface:
  * ``port``      -- oracle/liboracle.so (fdtd_oracle.c, the C restatement; ``*_omp`` = OpenMP build)
  * ``reference`` -- oracle/_ref/libref_openacc*.so, the UNMODIFIED /root/reference/openacc.cpp
                     compiled for the host (entry point ``Kernel_OpenACC``, openacc.cpp:61)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HALO = 4
WARMUP_STEPS = 5


class Geom(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("nxp", "nyp", "nzp", "x_m", "x_M", "y_m", "y_M", "z_m", "z_M")] + [
        (k, C.c_float) for k in ("dt", "h_x", "h_y", "h_z", "o_x", "o_y", "o_z")
    ]


class Dataobj(C.Structure):  # reference main.cpp:35-45
    _fields_ = [
        ("data", C.c_void_p),
        ("size", C.POINTER(C.c_int)),
        ("nbytes", C.c_ulong),
        ("npsize", C.c_void_p),
        ("dsize", C.c_void_p),
        ("hsize", C.c_void_p),
        ("hofs", C.c_void_p),
        ("oofs", C.c_void_p),
        ("dmap", C.c_void_p),
    ]


class Profiler(C.Structure):  # reference main.cpp:47-50
    _fields_ = [("section0", C.c_double), ("section1", C.c_double)]


def build(quiet: bool = True) -> None:
    """Run oracle/Makefile (also builds oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


_libs: dict = {}


def _lib(name: str):
    if name not in _libs:
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            build()
        _libs[name] = C.CDLL(path)
    return _libs[name]


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_openacc.so"))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def make_geom(nx, ny, nz, dt=1e-3, h=0.1, o=0.0, halo=HALO) -> Geom:
    h = (h, h, h) if np.isscalar(h) else h
    o = (o, o, o) if np.isscalar(o) else o
    return Geom(nx + 2 * halo, ny + 2 * halo, nz + 2 * halo, 0, nx - 1, 0, ny - 1, 0, nz - 1,
                dt, h[0], h[1], h[2], o[0], o[1], o[2])


# --------------------------------------------------------------------------- synthesis
def fill_ricker(T: int, S: int, dt: float = 1e-3) -> np.ndarray:
    """main.cpp:290-298."""
    out = np.empty((T, max(1, S)), np.float32)
    f = _lib("liboracle.so").oracle_fill_ricker
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
    f(_ptr(out), T, max(1, S), dt)
    return out


def fill_source_coords(S: int, nx: int, ny: int, nz: int, h=0.1) -> np.ndarray:
    """main.cpp:301-325."""
    out = np.zeros((max(1, S), 3), np.float32)
    f = _lib("liboracle.so").oracle_fill_source_coords
    f.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_float] * 3
    f(_ptr(out), S, nx, ny, nz, h, h, h)
    return out


def fill_dense(nx: int, ny: int, nz: int):
    """main.cpp:525-532 (level 2 zeroed explicitly)."""
    nxp, nyp, nzp = nx + 8, ny + 8, nz + 8
    u = np.empty((3, nxp, nyp, nzp), np.float32)
    m = np.empty((nxp, nyp, nzp), np.float32)
    f = _lib("liboracle.so").oracle_fill_dense
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    f(_ptr(u), _ptr(m), nxp * nyp * nzp)
    return u, m


def source_pos(coord: float, o: float, h: float):
    """openacc.cpp:125-131 for one axis -> (pos, frac)."""
    f = _lib("liboracle.so").oracle_source_pos
    f.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    pos, frac = C.c_int(), C.c_float()
    f(coord, o, h, C.byref(pos), C.byref(frac))
    return pos.value, np.float32(frac.value)


# --------------------------------------------------------------------------- the operator
def run(u, m, src=None, coords=None, *, dt=1e-3, h=0.1, o=0.0, time_m=0, time_M=None, p_src_m=0,
        p_src_M=None, impl="port", threads=1, extents=None):
    """Advance ``u`` (float32 [3, nxp, nyp, nzp], modified in place) from time_m to time_M inclusive.

    impl: "port" (C restatement) | "reference" (unmodified openacc.cpp on the host).
    threads > 1 selects the OpenMP build of either.  Returns (section0_s, section1_s) of the
    steps time >= time_m + 5, exactly like the reference's timers.
    extents: optional (x_m, x_M, y_m, y_M, z_m, z_M) override (default: full interior).
    """
    assert u.dtype == np.float32 and u.flags.c_contiguous and u.ndim == 4 and u.shape[0] == 3
    assert m.dtype == np.float32 and m.flags.c_contiguous and m.shape == u.shape[1:]
    nxp, nyp, nzp = u.shape[1:]
    g = make_geom(nxp - 8, nyp - 8, nzp - 8, dt, h, o)
    if extents is not None:
        g.x_m, g.x_M, g.y_m, g.y_M, g.z_m, g.z_M = extents
    has_src = src is not None and coords is not None and src.size > 0
    if has_src:
        src = _f32(src)
        coords = _f32(coords)
        assert src.ndim == 2 and coords.ndim == 2
        if p_src_M is None:
            p_src_M = coords.shape[0] - 1
        if time_M is None:
            time_M = src.shape[0] - 1
    else:
        p_src_M = -1
        assert time_M is not None
    if threads > 1:
        os.environ["OMP_NUM_THREADS"] = str(threads)

    if impl == "port":
        lib = _lib("liboracle_omp.so" if threads > 1 else "liboracle.so")
        f = lib.oracle_run
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Geom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        timers = (C.c_double * 2)(0.0, 0.0)
        rc = f(C.byref(g), _ptr(m), _ptr(u), _ptr(src) if has_src else None,
               src.shape[0] if has_src else 0, src.shape[1] if has_src else 1,
               _ptr(coords) if has_src else None, coords.shape[1] if has_src else 3,
               p_src_m, p_src_M, time_m, time_M, timers)
        assert rc == 0
        return timers[0], timers[1]

    if impl == "reference":
        lib = _lib(os.path.join("_ref", "libref_openacc_omp.so" if threads > 1 else "libref_openacc.so"))
        f = lib.Kernel_OpenACC
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Dataobj)] * 4 + [C.c_int] * 6 + [C.c_float] * 7 + [C.c_int] * 6 + [C.POINTER(Profiler)]

        def obj(arr, shape):
            sizes = (C.c_int * len(shape))(*shape)
            d = Dataobj()
            d.data = arr.ctypes.data if arr is not None and arr.size else None
            d.size = C.cast(sizes, C.POINTER(C.c_int))
            d.nbytes = int(np.prod(shape)) * 4
            d._keep = sizes
            return d

        if has_src:
            src_o, crd_o = obj(src, src.shape), obj(coords, coords.shape)
        else:  # main.cpp:537-545
            src_o, crd_o = obj(None, (0, 1)), obj(None, (2, 0))
        m_o, u_o = obj(m, m.shape), obj(u, u.shape)
        t = Profiler(0.0, 0.0)
        rc = f(C.byref(m_o), C.byref(src_o), C.byref(crd_o), C.byref(u_o), g.x_M, g.x_m, g.y_M, g.y_m,
               g.z_M, g.z_m, g.dt, g.h_x, g.h_y, g.h_z, g.o_x, g.o_y, g.o_z, p_src_M, p_src_m, time_M,
               time_m, -1, 1, C.byref(t))
        assert rc == 0
        return t.section0, t.section1

    raise ValueError(impl)


def fd_coeffs(space_order: int) -> np.ndarray:
    """Second-derivative weights c[0..R] of space order 2R as this repo defines them (fdtd_oracle.c)."""
    c = np.zeros(7, np.float32)
    f = _lib("liboracle.so").oracle_fd_coeffs
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_void_p]
    R = f(space_order, _ptr(c))
    if R < 0:
        raise ValueError("space_order must be 4, 6, 8, 10 or 12")
    return c[:R + 1]


def run_order(u, m, src=None, coords=None, *, space_order=4, rec_coords=None, dt=1e-3, h=0.1, o=0.0, time_m=0,
              time_M=None, p_src_m=0, p_src_M=None, threads=1):
    """The operator at space order 2R (halo = space_order cells, so u is [3, nx + 2*so, ...]) with optional receiver
    sampling.  Returns (section0_s, section1_s, rec) with rec float32 [T, nrec] or None.  Port only: the reference
    has no kernels beyond order 4 and no receivers; at order 4 without receivers this equals run()."""
    H = space_order
    assert u.dtype == np.float32 and u.flags.c_contiguous and u.ndim == 4 and u.shape[0] == 3
    assert m.dtype == np.float32 and m.flags.c_contiguous and m.shape == u.shape[1:]
    nxp, nyp, nzp = u.shape[1:]
    g = make_geom(nxp - 2 * H, nyp - 2 * H, nzp - 2 * H, dt, h, o, halo=H)
    has_src = src is not None and coords is not None and src.size > 0
    if has_src:
        src, coords = _f32(src), _f32(coords)
        p_src_M = coords.shape[0] - 1 if p_src_M is None else p_src_M
        time_M = src.shape[0] - 1 if time_M is None else time_M
    else:
        p_src_M = -1
        assert time_M is not None
    T = time_M - time_m + 1
    rec = None
    if rec_coords is not None and len(rec_coords):
        rec_coords = _f32(rec_coords)
        rec = np.zeros((T, rec_coords.shape[0]), np.float32)
    if threads > 1:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    f = _lib("liboracle_omp.so" if threads > 1 else "liboracle.so").oracle_run_order
    f.restype = C.c_int
    f.argtypes = [C.POINTER(Geom), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    timers = (C.c_double * 2)(0.0, 0.0)
    rc = f(C.byref(g), space_order, _ptr(m), _ptr(u), _ptr(src) if has_src else None, src.shape[0] if has_src else 0,
           src.shape[1] if has_src else 1, _ptr(coords) if has_src else None, coords.shape[1] if has_src else 3,
           p_src_m, p_src_M, time_m, time_M, _ptr(rec_coords) if rec is not None else None,
           rec.shape[1] if rec is not None else 0, rec_coords.shape[1] if rec is not None else 3,
           _ptr(rec) if rec is not None else None, timers)
    assert rc == 0
    return timers[0], timers[1], rec


def section0(u0, u1, m, *, dt=1e-3, h=0.1):
    """One Section0 application; returns the new level (halo cells zero)."""
    nxp, nyp, nzp = m.shape
    g = make_geom(nxp - 8, nyp - 8, nzp - 8, dt, h, 0.0)
    u2 = np.zeros_like(u0)
    f = _lib("liboracle.so").oracle_section0
    f.argtypes = [C.POINTER(Geom)] + [C.c_void_p] * 4
    f(C.byref(g), _ptr(m), _ptr(u0), _ptr(u1), _ptr(u2))
    return u2


# --------------------------------------------------------------------------- metrics
def rel_l2(a, b) -> float:
    """Relative L2 error of a against reference b, as main.cpp:589-593."""
    a64, b64 = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.sqrt(np.sum((a64 - b64) ** 2) / (np.sum(b64 ** 2) + 1e-30)))
