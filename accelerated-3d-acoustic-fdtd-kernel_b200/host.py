"""ctypes host mirror of include/fdtd_b200.h.

Part 1 mirrors the reference's operator boundary (main.cpp:35-80): ``Dataobj``/``Profiler`` structs
and ``Kernel_B200`` / ``Kernel_CUDA_Optimized`` with the reference's 24 arguments.  The Python
wrappers take numpy arrays where the C ABI takes ``dataobj*`` and build the descriptors the way
main.cpp:114-126,359-367 does.  Part 2 wraps the resident plan API.

No fallback: if libfdtd_b200.so is missing or a CUDA call fails, ``FdtdError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HALO = 4           # main.cpp:27-32
WARMUP_STEPS = 5   # openacc.cpp:5
IPC_BYTES = 160

_LIB_NAME = "libfdtd_b200.so"


class FdtdError(RuntimeError):
    """A libfdtd_b200 call returned a non-zero cudaError_t (or the library is unavailable)."""

    def __init__(self, what: str, code: int = -1):
        super().__init__(f"{what}: cudaError {code}" if code >= 0 else what)
        self.code = code


class Dataobj(C.Structure):  # main.cpp:35-45
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


class Profiler(C.Structure):  # main.cpp:47-50
    _fields_ = [("section0", C.c_double), ("section1", C.c_double)]


class Checksum(C.Structure):  # fdtd_b200_checksum
    _fields_ = [("bit_sum", C.c_ulonglong), ("pos_sum", C.c_ulonglong), ("nonzero", C.c_ulonglong),
                ("nonfinite", C.c_ulonglong), ("sum_sq", C.c_double), ("max_abs", C.c_float), ("reserved", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class Geometry(C.Structure):  # fdtd_b200_geometry
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("x_offset", C.c_int), ("nx_global", C.c_int),
        ("dt", C.c_float), ("h_x", C.c_float), ("h_y", C.c_float), ("h_z", C.c_float),
        ("o_x", C.c_float), ("o_y", C.c_float), ("o_z", C.c_float),
        ("deviceid", C.c_int),
    ]


def lib_path() -> str:
    return os.path.join(HERE, _LIB_NAME)


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into libfdtd_b200.so (in-tree).  nvcc cross-compiles without a GPU."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j8"]
    r = subprocess.run(cmd, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise FdtdError("building libfdtd_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return lib_path()


_KERNEL_ARGTYPES = ([C.POINTER(Dataobj)] * 4 + [C.c_int] * 6 + [C.c_float] * 7 + [C.c_int] * 6
                    + [C.POINTER(Profiler)])
_lib = None


def lib():
    """Load libfdtd_b200.so (raises FdtdError when it has not been built -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise FdtdError(f"{path} not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    for name in ("Kernel_B200", "Kernel_CUDA_Optimized"):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = i, _KERNEL_ARGTYPES
    L.FDTD_SetRuntimeConfig.restype, L.FDTD_SetRuntimeConfig.argtypes = None, [i, i, i]
    L.fdtd_b200_plan_create.restype, L.fdtd_b200_plan_create.argtypes = i, [C.POINTER(Geometry), C.POINTER(vp)]
    L.fdtd_b200_plan_create_order.restype, L.fdtd_b200_plan_create_order.argtypes = i, [C.POINTER(Geometry), i, C.POINTER(vp)]
    L.fdtd_b200_plan_set_receivers.restype, L.fdtd_b200_plan_set_receivers.argtypes = i, [vp, vp, i, i]
    L.fdtd_b200_plan_download_receivers.restype = i
    L.fdtd_b200_plan_download_receivers.argtypes = [vp, vp, C.POINTER(i), C.POINTER(i)]
    L.fdtd_b200_plan_destroy.restype, L.fdtd_b200_plan_destroy.argtypes = i, [vp]
    L.fdtd_b200_plan_u.restype, L.fdtd_b200_plan_u.argtypes = vp, [vp]
    L.fdtd_b200_plan_m.restype, L.fdtd_b200_plan_m.argtypes = vp, [vp]
    L.fdtd_b200_plan_level.restype, L.fdtd_b200_plan_level.argtypes = vp, [vp, i]
    L.fdtd_b200_plan_probe_fuse.restype, L.fdtd_b200_plan_probe_fuse.argtypes = i, [vp, C.POINTER(i)]
    L.fdtd_b200_plan_level_elems.restype, L.fdtd_b200_plan_level_elems.argtypes = C.c_size_t, [vp]
    L.fdtd_b200_plan_upload.restype, L.fdtd_b200_plan_upload.argtypes = i, [vp, vp, vp]
    L.fdtd_b200_plan_download.restype, L.fdtd_b200_plan_download.argtypes = i, [vp, vp]
    L.fdtd_b200_plan_download_window.restype, L.fdtd_b200_plan_download_window.argtypes = i, [vp] + [i] * 7 + [vp]
    L.fdtd_b200_plan_checksum.restype, L.fdtd_b200_plan_checksum.argtypes = i, [vp] + [i] * 7 + [C.POINTER(Checksum)]
    L.fdtd_b200_plan_fill.restype, L.fdtd_b200_plan_fill.argtypes = i, [vp, f, f]
    L.fdtd_b200_plan_fill_dense.restype, L.fdtd_b200_plan_fill_dense.argtypes = i, [vp]
    L.fdtd_b200_plan_set_sources.restype, L.fdtd_b200_plan_set_sources.argtypes = i, [vp, vp, i, i, vp, i, i, i, i]
    L.fdtd_b200_plan_run.restype, L.fdtd_b200_plan_run.argtypes = i, [vp, i, i, C.POINTER(Profiler)]
    L.fdtd_b200_plan_run_staged.restype, L.fdtd_b200_plan_run_staged.argtypes = i, [vp, vp, vp, i, i, C.POINTER(Profiler)]
    L.fdtd_b200_plan_last_launches.restype, L.fdtd_b200_plan_last_launches.argtypes = C.c_long, [vp]
    L.fdtd_b200_plan_last_kernel_seconds.restype, L.fdtd_b200_plan_last_kernel_seconds.argtypes = d, [vp]
    L.fdtd_b200_plan_set_option.restype, L.fdtd_b200_plan_set_option.argtypes = i, [vp, C.c_char_p, i]
    L.fdtd_b200_plan_get_option.restype, L.fdtd_b200_plan_get_option.argtypes = i, [vp, C.c_char_p, C.POINTER(i)]
    L.fdtd_b200_plan_ipc_export.restype, L.fdtd_b200_plan_ipc_export.argtypes = i, [vp, vp]
    L.fdtd_b200_plan_ipc_attach.restype, L.fdtd_b200_plan_ipc_attach.argtypes = i, [vp, i, vp]
    L.fdtd_b200_plan_attach_local.restype, L.fdtd_b200_plan_attach_local.argtypes = i, [vp, i, vp]
    L.fdtd_b200_run_slabs.restype, L.fdtd_b200_run_slabs.argtypes = i, [C.POINTER(vp), i, i, i, C.POINTER(Profiler)]
    L.fdtd_b200_source_table.restype = i
    L.fdtd_b200_source_table.argtypes = [C.POINTER(f)] * 3 + [C.POINTER(i)] * 2 + [C.POINTER(i), C.POINTER(f),
                                                                                    C.POINTER(f), C.POINTER(i)]
    L.fdtd_b200_slab_source_cells.restype = i
    L.fdtd_b200_slab_source_cells.argtypes = [C.POINTER(Geometry), vp, i, i, i, i, i, vp, C.POINTER(i), C.POINTER(i), i, vp, vp,
                                              C.POINTER(i), vp]
    L.fdtd_b200_slab_source_cells2.restype = i
    L.fdtd_b200_slab_source_cells2.argtypes = [C.POINTER(Geometry), vp, i, i, i, i, i, vp, C.POINTER(i), i, vp, vp, C.POINTER(i),
                                               C.POINTER(i)]
    L.fdtd_b200_fill_ricker.restype, L.fdtd_b200_fill_ricker.argtypes = None, [vp, i, i, f]
    L.fdtd_b200_fill_source_coords.restype, L.fdtd_b200_fill_source_coords.argtypes = None, [vp, i, i, i, i, f, f, f]
    L.fdtd_b200_write_benchmark_csv.restype = i
    L.fdtd_b200_write_benchmark_csv.argtypes = [C.c_char_p, C.c_char_p] + [d] * 17 + [i] * 6
    L.fdtd_b200_version.restype, L.fdtd_b200_version.argtypes = C.c_char_p, []
    _lib = L
    return L


def exported_symbols():
    """Every entry point include/fdtd_b200.h declares (checked against the .so by the CPU tests)."""
    return [
        "Kernel_CUDA_Optimized", "Kernel_B200", "FDTD_SetRuntimeConfig",
        "fdtd_b200_plan_create", "fdtd_b200_plan_create_order", "fdtd_b200_plan_set_receivers", "fdtd_b200_plan_download_receivers",
        "fdtd_b200_plan_destroy", "fdtd_b200_plan_u", "fdtd_b200_plan_m",
        "fdtd_b200_plan_level", "fdtd_b200_plan_probe_fuse", "fdtd_b200_plan_level_elems", "fdtd_b200_plan_upload", "fdtd_b200_plan_download", "fdtd_b200_plan_download_window", "fdtd_b200_plan_checksum", "fdtd_b200_plan_fill",
        "fdtd_b200_plan_fill_dense", "fdtd_b200_plan_set_sources", "fdtd_b200_plan_run", "fdtd_b200_plan_run_staged",
        "fdtd_b200_plan_last_launches", "fdtd_b200_plan_last_kernel_seconds", "fdtd_b200_plan_set_option",
        "fdtd_b200_plan_get_option", "fdtd_b200_plan_ipc_export", "fdtd_b200_plan_ipc_attach",
        "fdtd_b200_plan_attach_local", "fdtd_b200_run_slabs", "fdtd_b200_source_table", "fdtd_b200_slab_source_cells", "fdtd_b200_slab_source_cells2", "fdtd_b200_fill_ricker",
        "fdtd_b200_fill_source_coords", "fdtd_b200_write_benchmark_csv", "fdtd_b200_version",
    ]


def _check(rc: int, what: str):
    if rc != 0:
        raise FdtdError(what, rc)


# --------------------------------------------------------------------------- Part 1: reference ABI
def make_dataobj(arr, shape=None) -> Dataobj:
    """initialize_dataobj of main.cpp:114-126: data, size[], nbytes; everything else null."""
    shape = tuple(arr.shape) if shape is None else tuple(shape)
    sizes = (C.c_int * len(shape))(*shape)
    d = Dataobj()
    d.data = arr.ctypes.data if arr is not None and arr.size else None
    d.size = C.cast(sizes, C.POINTER(C.c_int))
    d.nbytes = int(np.prod(shape)) * 4 if arr is not None else 0
    d._keep = (sizes, arr)
    return d


def _kernel(name, m, src, src_coords, u, x_M, x_m, y_M, y_m, z_M, z_m, dt, h_x, h_y, h_z, o_x, o_y, o_z,
            p_src_M, p_src_m, time_M, time_m, deviceid, devicerm, timers):
    for a, nm in ((u, "u"), (m, "m")):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous):
            raise TypeError(f"{nm} must be a C-contiguous float32 numpy array")
    m_o, u_o = make_dataobj(m), make_dataobj(u)
    if src is None or src_coords is None:  # the correctness test's "no sources" descriptors, main.cpp:537-545
        src_o, crd_o = make_dataobj(None, (0, 1)), make_dataobj(None, (2, 0))
    else:
        src = np.ascontiguousarray(src, np.float32)
        src_coords = np.ascontiguousarray(src_coords, np.float32)
        src_o, crd_o = make_dataobj(src), make_dataobj(src_coords)
    t = timers if timers is not None else Profiler(0.0, 0.0)
    rc = getattr(lib(), name)(C.byref(m_o), C.byref(src_o), C.byref(crd_o), C.byref(u_o), x_M, x_m, y_M, y_m, z_M,
                              z_m, dt, h_x, h_y, h_z, o_x, o_y, o_z, p_src_M, p_src_m, time_M, time_m, deviceid,
                              devicerm, C.byref(t))
    return rc


def Kernel_B200(m_vec, src_vec, src_coords_vec, u_vec, x_M, x_m, y_M, y_m, z_M, z_m, dt, h_x, h_y, h_z, o_x, o_y,
                o_z, p_src_M, p_src_m, time_M, time_m, deviceid=0, devicerm=1, timers=None) -> int:
    """The operator, argument for argument as main.cpp:53-58 (max before min); returns the C int."""
    return _kernel("Kernel_B200", m_vec, src_vec, src_coords_vec, u_vec, x_M, x_m, y_M, y_m, z_M, z_m, dt, h_x, h_y,
                   h_z, o_x, o_y, o_z, p_src_M, p_src_m, time_M, time_m, deviceid, devicerm, timers)


def Kernel_CUDA_Optimized(m_vec, src_vec, src_coords_vec, u_vec, x_M, x_m, y_M, y_m, z_M, z_m, dt, h_x, h_y, h_z,
                          o_x, o_y, o_z, p_src_M, p_src_m, time_M, time_m, deviceid=0, devicerm=1, timers=None) -> int:
    """Drop-in name of the reference's optimized entry point (main.cpp:67-72)."""
    return _kernel("Kernel_CUDA_Optimized", m_vec, src_vec, src_coords_vec, u_vec, x_M, x_m, y_M, y_m, z_M, z_m, dt,
                   h_x, h_y, h_z, o_x, o_y, o_z, p_src_M, p_src_m, time_M, time_m, deviceid, devicerm, timers)


def FDTD_SetRuntimeConfig(use_tc: int, t_fuse: int, nfields: int) -> None:
    lib().FDTD_SetRuntimeConfig(use_tc, t_fuse, nfields)


# --------------------------------------------------------------------------- host-only helpers
def fill_ricker(T: int, S: int, dt: float = 1e-3) -> np.ndarray:
    out = np.empty((T, max(1, S)), np.float32)
    lib().fdtd_b200_fill_ricker(out.ctypes.data, T, max(1, S), dt)
    return out


def fill_source_coords(S: int, nx: int, ny: int, nz: int, h=(0.1, 0.1, 0.1)) -> np.ndarray:
    out = np.zeros((max(1, S), 3), np.float32)
    lib().fdtd_b200_fill_source_coords(out.ctypes.data, S, nx, ny, nz, h[0], h[1], h[2])
    return out


def source_table(coord, o, h, lo, hi):
    """(pos[3], frac[3], w[8], in_range[8]) of one source, openacc.cpp:125-134 (host, IEEE fp32)."""
    f3, i3 = C.c_float * 3, C.c_int * 3
    pos, frac, w, inr = i3(), f3(), (C.c_float * 8)(), (C.c_int * 8)()
    _check(lib().fdtd_b200_source_table(f3(*coord), f3(*o), f3(*h), i3(*lo), i3(*hi), pos, frac, w, inr),
           "fdtd_b200_source_table")
    return (np.array(pos[:], np.int32), np.array(frac[:], np.float32), np.array(w[:], np.float32),
            np.array(inr[:], np.int32))


def slab_source_cells(geom: Geometry, coords, p_src_m=0, p_src_M=None):
    """Host-only scatter table of one slab -> (cells[n,5], ncells_int, contrib_p, contrib_w, base_idx)."""
    coords = np.ascontiguousarray(coords, np.float32)
    p_src_M = coords.shape[0] - 1 if p_src_M is None else p_src_M
    nsrc = max(0, p_src_M - p_src_m + 1)
    cells = np.zeros((8 * nsrc + 1, 5), np.int32)
    cp = np.zeros(8 * nsrc + 1, np.int32)
    cw = np.zeros(8 * nsrc + 1, np.float32)
    base = np.full(p_src_M + 2, -1, np.int64)
    n_int, n_all, n_c = C.c_int(), C.c_int(), C.c_int()
    _check(lib().fdtd_b200_slab_source_cells(C.byref(geom), coords.ctypes.data, coords.shape[0], coords.shape[1], p_src_m,
                                             p_src_M, cells.shape[0], cells.ctypes.data, C.byref(n_int), C.byref(n_all),
                                             cp.shape[0], cp.ctypes.data, cw.ctypes.data, C.byref(n_c), base.ctypes.data),
           "fdtd_b200_slab_source_cells")
    return cells[:n_all.value], n_int.value, cp[:n_c.value], cw[:n_c.value], base[:p_src_M + 1]


def slab_source_cells2(geom: Geometry, coords, p_src_m=0, p_src_M=None):
    """Host-only table of the two-step launches -> (cells[n,5], contrib_p, contrib_w, halo_global)."""
    coords = np.ascontiguousarray(coords, np.float32)
    p_src_M = coords.shape[0] - 1 if p_src_M is None else p_src_M
    nsrc = max(0, p_src_M - p_src_m + 1)
    cells = np.zeros((16 * nsrc + 1, 5), np.int32)
    cp = np.zeros(16 * nsrc + 1, np.int32)
    cw = np.zeros(16 * nsrc + 1, np.float32)
    n_all, n_c, hg = C.c_int(), C.c_int(), C.c_int()
    _check(lib().fdtd_b200_slab_source_cells2(C.byref(geom), coords.ctypes.data, coords.shape[0], coords.shape[1], p_src_m,
                                              p_src_M, cells.shape[0], cells.ctypes.data, C.byref(n_all), cp.shape[0],
                                              cp.ctypes.data, cw.ctypes.data, C.byref(n_c), C.byref(hg)),
           "fdtd_b200_slab_source_cells2")
    return cells[:n_all.value], cp[:n_c.value], cw[:n_c.value], bool(hg.value)


def write_benchmark_csv(filename, method, total, s0, s1, device, overhead, gflops, gbps, peak_fp32_gf, peak_bw_gbs,
                        ai, nx, ny, nz, timesteps, nsrc, stencil_order=4):
    """Append one row in the reference's 24-column schema; each of total..gbps is a (mean, std) pair."""
    args = [filename.encode(), method.encode()]
    for pair in (total, s0, s1, device, overhead, gflops, gbps):
        args += [float(pair[0]), float(pair[1])]
    args += [float(peak_fp32_gf), float(peak_bw_gbs), float(ai), nx, ny, nz, timesteps, nsrc, stencil_order]
    _check(lib().fdtd_b200_write_benchmark_csv(*args), "fdtd_b200_write_benchmark_csv")


# --------------------------------------------------------------------------- Part 2: resident plans
class Plan:
    """One x-slab resident on one GPU (the whole grid when x_offset = 0 and nx_global = nx)."""

    def __init__(self, nx, ny, nz, *, dt=1e-3, h=(0.1, 0.1, 0.1), o=(0.0, 0.0, 0.0), x_offset=0, nx_global=None,
                 deviceid=-1, space_order=4):
        h = (h, h, h) if np.isscalar(h) else h
        o = (o, o, o) if np.isscalar(o) else o
        self.geom = Geometry(nx, ny, nz, x_offset, nx if nx_global is None else nx_global, dt, h[0], h[1], h[2],
                             o[0], o[1], o[2], deviceid)
        self.halo = space_order  # main.cpp:27-32: HALO == STENCIL_ORDER cells per side
        self.shape = (3, nx + 2 * self.halo, ny + 2 * self.halo, nz + 2 * self.halo)
        self._h = C.c_void_p()
        if space_order == HALO:
            _check(lib().fdtd_b200_plan_create(C.byref(self.geom), C.byref(self._h)), "fdtd_b200_plan_create")
        else:
            _check(lib().fdtd_b200_plan_create_order(C.byref(self.geom), space_order, C.byref(self._h)),
                   "fdtd_b200_plan_create_order")
        self._nrec = 0

    def close(self):
        if self._h:
            lib().fdtd_b200_plan_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def u_ptr(self) -> int:
        return lib().fdtd_b200_plan_u(self._h)

    @property
    def m_ptr(self) -> int:
        return lib().fdtd_b200_plan_m(self._h)

    def level_ptr(self, ring_level: int) -> int:
        """Device pointer of ring level 0..2 (two-step passes rotate a spare device level through the ring)."""
        return lib().fdtd_b200_plan_level(self._h, ring_level)

    def probe_fuse(self) -> int:
        """Time steps per pass (1 or 2) this slab could run now; linked slabs must agree on the minimum."""
        v = C.c_int()
        _check(lib().fdtd_b200_plan_probe_fuse(self._h, C.byref(v)), "fdtd_b200_plan_probe_fuse")
        return v.value

    @property
    def level_elems(self) -> int:
        return lib().fdtd_b200_plan_level_elems(self._h)

    def upload(self, u=None, m=None):
        for a, shp in ((u, self.shape), (m, self.shape[1:])):
            if a is not None and not (a.dtype == np.float32 and a.flags.c_contiguous and a.shape == shp):
                raise TypeError("upload expects C-contiguous float32 arrays of the padded shape")
        _check(lib().fdtd_b200_plan_upload(self._h, u.ctypes.data if u is not None else None,
                                           m.ctypes.data if m is not None else None), "fdtd_b200_plan_upload")

    def download(self, out=None) -> np.ndarray:
        out = np.empty(self.shape, np.float32) if out is None else out
        _check(lib().fdtd_b200_plan_download(self._h, out.ctypes.data), "fdtd_b200_plan_download")
        return out

    def _window(self, window):
        """(x0, x1, y0, y1, z0, z1) in padded local coordinates; None = the whole padded level."""
        if window is None:
            return (0, self.shape[1], 0, self.shape[2], 0, self.shape[3])
        w = tuple(int(v) for v in window)
        if len(w) != 6:
            raise ValueError("window = (x0, x1, y0, y1, z0, z1)")
        return w

    def interior(self):
        """The window of the cells Section0 updates."""
        H = self.halo
        return (H, self.shape[1] - H, H, self.shape[2] - H, H, self.shape[3] - H)

    def download_window(self, ring_level: int, window) -> np.ndarray:
        """One ring level on [x0,x1) x [y0,y1) x [z0,z1) (padded local coordinates) as a dense array."""
        w = self._window(window)
        out = np.empty((w[1] - w[0], w[3] - w[2], w[5] - w[4]), np.float32)
        _check(lib().fdtd_b200_plan_download_window(self._h, ring_level, *w, out.ctypes.data),
               "fdtd_b200_plan_download_window")
        return out

    def checksum(self, ring_level: int, window=None) -> dict:
        """Device-side checksum of a window: order-independent integer sums (they add up over slabs), max|u|, sum u^2."""
        c = Checksum()
        _check(lib().fdtd_b200_plan_checksum(self._h, ring_level, *self._window(window), C.byref(c)),
               "fdtd_b200_plan_checksum")
        return c.as_dict()

    def fill(self, u_value=0.0, m_value=1.5):
        _check(lib().fdtd_b200_plan_fill(self._h, u_value, m_value), "fdtd_b200_plan_fill")

    def fill_dense(self):
        _check(lib().fdtd_b200_plan_fill_dense(self._h), "fdtd_b200_plan_fill_dense")

    def set_sources(self, src, coords, p_src_m=0, p_src_M=None):
        src = np.ascontiguousarray(src, np.float32)
        coords = np.ascontiguousarray(coords, np.float32)
        p_src_M = coords.shape[0] - 1 if p_src_M is None else p_src_M
        _check(lib().fdtd_b200_plan_set_sources(self._h, src.ctypes.data, src.shape[0], src.shape[1],
                                                coords.ctypes.data, coords.shape[0], coords.shape[1], p_src_m,
                                                p_src_M), "fdtd_b200_plan_set_sources")

    def set_receivers(self, coords):
        """Receiver positions [nrec, >=3] in global physical coordinates; every run then records rec[time][r]."""
        coords = np.ascontiguousarray(coords, np.float32).reshape(-1, 3) if coords is not None and len(coords) else None
        self._nrec = 0 if coords is None else coords.shape[0]
        _check(lib().fdtd_b200_plan_set_receivers(self._h, coords.ctypes.data if coords is not None else None, self._nrec,
                                                  3), "fdtd_b200_plan_set_receivers")

    def receivers(self):
        """(rec [rows, nrec] float32, owned [nrec] bool) of the last run; owned marks the receivers THIS slab sampled."""
        rows = C.c_int()
        _check(lib().fdtd_b200_plan_download_receivers(self._h, None, None, C.byref(rows)), "fdtd_b200_plan_download_receivers")
        rec = np.zeros((rows.value, self._nrec), np.float32)
        owned = (C.c_int * max(1, self._nrec))()
        _check(lib().fdtd_b200_plan_download_receivers(self._h, rec.ctypes.data if rec.size else None, owned, C.byref(rows)),
               "fdtd_b200_plan_download_receivers")
        return rec, np.array(owned[:self._nrec], bool)

    def run(self, time_m: int, time_M: int) -> Profiler:
        t = Profiler(0.0, 0.0)
        _check(lib().fdtd_b200_plan_run(self._h, time_m, time_M, C.byref(t)), "fdtd_b200_plan_run")
        return t

    def run_staged(self, u: np.ndarray, m: np.ndarray, time_m: int, time_M: int):
        """upload + run + download as one pipeline (u is updated in place).  Returns the Profiler, or None when
        the library asks for the three-phase path (cudaErrorNotSupported)."""
        for a, shp in ((u, self.shape), (m, self.shape[1:])):
            if not (a.dtype == np.float32 and a.flags.c_contiguous and a.shape == shp):
                raise TypeError("run_staged expects C-contiguous float32 arrays of the padded shape")
        t = Profiler(0.0, 0.0)
        rc = lib().fdtd_b200_plan_run_staged(self._h, u.ctypes.data, m.ctypes.data, time_m, time_M, C.byref(t))
        if rc == 801:
            return None
        _check(rc, "fdtd_b200_plan_run_staged")
        return t

    def set_option(self, key: str, value: int):
        _check(lib().fdtd_b200_plan_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def get_option(self, key: str) -> int:
        v = C.c_int()
        _check(lib().fdtd_b200_plan_get_option(self._h, key.encode(), C.byref(v)), f"get_option({key})")
        return v.value

    @property
    def last_launches(self) -> int:
        return lib().fdtd_b200_plan_last_launches(self._h)

    @property
    def last_kernel_seconds(self) -> float:
        return lib().fdtd_b200_plan_last_kernel_seconds(self._h)

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_BYTES)
        _check(lib().fdtd_b200_plan_ipc_export(self._h, buf), "fdtd_b200_plan_ipc_export")
        return buf.raw

    def ipc_attach(self, side: int, blob: bytes):
        _check(lib().fdtd_b200_plan_ipc_attach(self._h, side, C.create_string_buffer(blob, IPC_BYTES)),
               "fdtd_b200_plan_ipc_attach")

    def attach_local(self, side: int, other: "Plan"):
        _check(lib().fdtd_b200_plan_attach_local(self._h, side, other._h), "fdtd_b200_plan_attach_local")


def run_slabs(plans, time_m: int, time_M: int) -> Profiler:
    """Advance several slabs driven by this process in lock step (fdtd_b200_run_slabs)."""
    arr = (C.c_void_p * len(plans))(*[p.handle for p in plans])
    t = Profiler(0.0, 0.0)
    _check(lib().fdtd_b200_run_slabs(arr, len(plans), time_m, time_M, C.byref(t)), "fdtd_b200_run_slabs")
    return t
