"""ctypes binding of libns3d.so -- one Python function per entry point of include/ns3d.h.

This is the executable stand-in for the Julia ``ccall`` shim (julia/NS3DNative.jl): same
symbols, same argument order.  There is no fallback of any kind: if the shared object is
missing or a call fails, an exception is raised (``NS3DError``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

PARITY, FAST, FASTEST = 0, 1, 2
STREAM_COMPUTE, STREAM_H2D, STREAM_D2H = 0, 1, 2
VARIANT_M, VARIANT_G = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class NS3DError(RuntimeError):
    pass


class PtParams(C.Structure):
    """``ns3d_pt_params`` (include/ns3d.h)."""
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("variant", C.c_int),
        ("rho", C.c_double), ("dt", C.c_double), ("dtau", C.c_double), ("damp", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("eps_it", C.c_double), ("err_num", C.c_double), ("err_den", C.c_double),
        ("niter", C.c_int), ("nchk", C.c_int), ("outlet_guard", C.c_int),
        ("outlet_val", C.c_double), ("g", C.c_double),
        ("zchunk", C.c_int), ("reserved", C.c_int),
    ]


FIELD_NAMES = ["Pr", "dPrdtau", "C", "C_o", "txx", "tyy", "tzz", "txy", "txz", "tyz",
               "Vx", "Vy", "Vz", "Vx_o", "Vy_o", "Vz_o", "divV", "Rp"]


class Fields(C.Structure):
    """``ns3d_fields``: 18 device pointers in the reference's allocation order (M:343-360)."""
    _fields_ = [(n, C.c_void_p) for n in FIELD_NAMES]


class StepParams(C.Structure):
    """``ns3d_step_params``."""
    _fields_ = [
        ("pt", PtParams),
        ("mu", C.c_double), ("vin", C.c_double),
        ("a2", C.c_double), ("b2", C.c_double), ("ox", C.c_double), ("oy", C.c_double),
        ("sinb", C.c_double), ("cosb", C.c_double),
        ("xco_g", C.c_double), ("yco_g", C.c_double), ("lx", C.c_double), ("ly", C.c_double),
        ("inlet_guard", C.c_int), ("reserved", C.c_int),
    ]


# name -> (restype, argtypes); the single source of truth for the symbols the header declares.
_D, _I, _P, _Z = C.c_double, C.c_int, C.c_void_p, C.c_size_t
SIGNATURES = {
    "ns3d_create": (_I, [_I, C.POINTER(_P)]),
    "ns3d_destroy": (_I, [_P]),
    "ns3d_version": (C.c_char_p, []),
    "ns3d_last_error": (C.c_char_p, [_P]),
    "ns3d_set_mode": (_I, [_P, _I]),
    "ns3d_get_mode": (_I, [_P]),
    "ns3d_set_option": (_I, [_P, C.c_char_p, _I]),
    "ns3d_sync": (_I, [_P]),
    "ns3d_launch_count": (C.c_longlong, [_P]),
    "ns3d_stream": (_P, [_P]),
    "ns3d_box_d2h": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I]),
    "ns3d_gather_box": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, c_int_p, _P, _I]),
    "ns3d_zeros": (_I, [_P, _I, _I, _I, C.POINTER(_P)]),
    "ns3d_free": (_I, [_P, _P]),
    "ns3d_h2d": (_I, [_P, _P, _P, _Z]),
    "ns3d_d2h": (_I, [_P, _P, _P, _Z]),
    "ns3d_copy": (_I, [_P, _P, _P, _Z]),
    "ns3d_fill_profile_z": (_I, [_P, _P, _I, _I, _I, c_double_p]),
    "ns3d_fill_profile_zy": (_I, [_P, _P, _I, _I, _I, c_double_p, c_double_p, c_double_p]),
    "ns3d_fill_plane_x": (_I, [_P, _P, _I, _I, _I, _I, _D]),
    "ns3d_h2d_async": (_I, [_P, _P, _P, _Z]),
    "ns3d_d2h_async": (_I, [_P, _P, _P, _Z]),
    "ns3d_stream_wait": (_I, [_P, _I, _I]),
    "ns3d_stream_sync": (_I, [_P, _I]),
    "ns3d_fill": (_I, [_P, _P, _D, _Z]),
    "ns3d_bytes_allocated": (_Z, [_P]),
    "ns3d_update_tau": (_I, [_P] + [_P] * 9 + [_D] * 4 + [_I] * 3),
    "ns3d_predict_V": (_I, [_P] + [_P] * 9 + [_D] * 6 + [_I] * 3),
    "ns3d_update_divV": (_I, [_P] + [_P] * 4 + [_D] * 3 + [_I] * 3),
    "ns3d_update_dPrdtau": (_I, [_P] + [_P] * 3 + [_D] * 7 + [_I] * 3),
    "ns3d_update_Pr": (_I, [_P, _P, _P, _D, _I, _I, _I]),
    "ns3d_compute_res": (_I, [_P] + [_P] * 3 + [_D] * 5 + [_I] * 3),
    "ns3d_max_abs": (_I, [_P, _P, _Z, c_double_p]),
    "ns3d_correct_V": (_I, [_P] + [_P] * 4 + [_D] * 5 + [_I] * 3),
    "ns3d_bc_x": (_I, [_P, _P, _I, _I, _I]),
    "ns3d_bc_y": (_I, [_P, _P, _I, _I, _I]),
    "ns3d_bc_z": (_I, [_P, _P, _I, _I, _I]),
    "ns3d_bc_x_Vx": (_I, [_P, _P, _D, _I, _I, _I]),
    "ns3d_bc_x_Pr": (_I, [_P, _P, _D, _I, _I, _I]),
    "ns3d_bc_zV": (_I, [_P, _P, _I, _I, _I]),
    "ns3d_bc_xhydstatic": (_I, [_P, _P, _D, _I, _D, _D, _I, _I, _I]),
    "ns3d_set_bc_Vel_M": (_I, [_P, _P, _P, _P, _I, _D, _I, _I, _I]),
    "ns3d_set_bc_Vel_G": (_I, [_P, _P, _P, _P, _I, _I, _I]),
    "ns3d_set_bc_Pr_M": (_I, [_P, _P, _I, _D, _I, _I, _I]),
    "ns3d_set_bc_Pr_G": (_I, [_P, _P, _D, _I, _D, _D, _I, _I, _I]),
    "ns3d_advect": (_I, [_P] + [_P] * 8 + [_D] * 4 + [_I] * 3),
    "ns3d_set_cylinder_M": (_I, [_P] + [_P] * 4 + [_D] * 10 + [_I] * 3),
    "ns3d_set_cylinder_G": (_I, [_P] + [_P] * 4 + [_D] * 10 + [_I] * 3),
    "ns3d_comm_unique_id": (_I, [C.c_char_p]),
    "ns3d_comm_init": (_I, [_P, _I, _I, C.c_char_p]),
    "ns3d_comm_rank": (_I, [_P]),
    "ns3d_comm_size": (_I, [_P]),
    "ns3d_update_halo": (_I, [_P, C.POINTER(_P), c_int_p, c_int_p, c_int_p, _I, _I]),
    "ns3d_allreduce_max": (_I, [_P, c_double_p]),
    "ns3d_pt_solve": (_I, [_P, _P, _P, _P, C.POINTER(PtParams), c_int_p, c_double_p, _I, c_int_p]),
    "ns3d_pt_iterate": (_I, [_P, _P, _P, _P, C.POINTER(PtParams), _I]),
    "ns3d_pt_describe": (_I, [_P, C.POINTER(PtParams), C.c_char_p, _I, c_int_p]),
    "ns3d_step": (_I, [_P, C.POINTER(Fields), C.POINTER(StepParams), c_int_p, c_double_p, _I, c_int_p]),
    "ns3d_predictor": (_I, [_P, C.POINTER(Fields), C.POINTER(StepParams)]),
    "ns3d_corrector": (_I, [_P, C.POINTER(Fields), C.POINTER(StepParams)]),
    "ns3d_advect_swap": (_I, [_P, C.POINTER(Fields), C.POINTER(StepParams)]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libns3d.so and type every symbol of include/ns3d.h.  Raises if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001
            if not os.path.exists(path):
                raise NS3DError(f"libns3d.so is missing and cannot be built: {exc}") from exc
    if not os.path.exists(path):
        raise NS3DError(f"{path} not found: build it with `python -m navierstokes3d_b200.build` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DeviceArray:
    """A dense column-major float64 device array of the reference's shape, owned by a Context."""

    __slots__ = ("ctx", "ptr", "shape")

    def __init__(self, ctx: "Context", ptr: int, shape):
        self.ctx, self.ptr, self.shape = ctx, ptr, tuple(shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def to_host(self) -> np.ndarray:
        """``Array(A)`` (M:399)."""
        out = np.empty(self.shape, dtype=np.float64, order="F")
        self.ctx.d2h(out, self)
        return out

    def set(self, host: np.ndarray) -> "DeviceArray":
        """``A = Data.Array(host)`` (M:370)."""
        self.ctx.h2d(self, host)
        return self


def _ptr(a) -> int:
    return a.ptr if isinstance(a, DeviceArray) else (0 if a is None else int(a))


class Context:
    """``ns3d_ctx``: one GPU, one stream, its field allocator and (optionally) a z-slab communicator."""

    def __init__(self, device: int = 0, mode: int = PARITY):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ns3d_create(device, C.byref(h))
        if rc != 0:
            raise NS3DError(f"ns3d_create({device}) failed ({rc}): {self.lib.ns3d_last_error(None).decode()}")
        self.h = h
        self.device = device
        self.set_mode(mode)

    # -- plumbing ------------------------------------------------------------------------------
    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise NS3DError(f"{what} failed ({rc}): {self.lib.ns3d_last_error(self.h).decode()}")

    def call(self, name: str, *args):
        """Call entry point ``name`` with the context prepended; DeviceArrays become pointers."""
        conv = [_ptr(a) if isinstance(a, DeviceArray) or a is None else a for a in args]
        self._ck(getattr(self.lib, name)(self.h, *conv), name)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ns3d_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def set_mode(self, mode: int):
        self._ck(self.lib.ns3d_set_mode(self.h, mode), "ns3d_set_mode")

    def set_option(self, name: str, value: int):
        self._ck(self.lib.ns3d_set_option(self.h, name.encode(), value), "ns3d_set_option")

    @property
    def mode(self) -> int:
        return self.lib.ns3d_get_mode(self.h)

    def sync(self):
        self._ck(self.lib.ns3d_sync(self.h), "ns3d_sync")

    @property
    def launch_count(self) -> int:
        return int(self.lib.ns3d_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.ns3d_stream(self.h) or 0)

    @property
    def bytes_allocated(self) -> int:
        return int(self.lib.ns3d_bytes_allocated(self.h))

    # -- allocator -----------------------------------------------------------------------------
    def zeros(self, *shape) -> DeviceArray:
        """``@zeros(sx,sy,sz)`` (M:343-360)."""
        p = C.c_void_p()
        self._ck(self.lib.ns3d_zeros(self.h, shape[0], shape[1], shape[2], C.byref(p)), "ns3d_zeros")
        return DeviceArray(self, p.value, shape)

    def free(self, a: DeviceArray):
        self._ck(self.lib.ns3d_free(self.h, a.ptr), "ns3d_free")
        a.ptr = 0

    def h2d(self, dst: DeviceArray, host: np.ndarray):
        host = np.asfortranarray(host, dtype=np.float64)
        if host.shape != dst.shape:
            raise NS3DError(f"h2d: shape {host.shape} != {dst.shape}")
        self._ck(self.lib.ns3d_h2d(self.h, dst.ptr, host.ctypes.data, host.size), "ns3d_h2d")

    def h2d_raw(self, dst_ptr: int, host_ptr: int, count: int):
        self._ck(self.lib.ns3d_h2d(self.h, dst_ptr, host_ptr, count), "ns3d_h2d")

    def d2h(self, host: np.ndarray, src: DeviceArray):
        assert host.dtype == np.float64 and host.flags.f_contiguous and host.shape == src.shape
        self._ck(self.lib.ns3d_d2h(self.h, host.ctypes.data, src.ptr, host.size), "ns3d_d2h")

    def box(self, src: DeviceArray, xr=None, yr=None, zr=None, dtype=np.float64) -> np.ndarray:
        """The box ``src[xr[0]:xr[1], yr[0]:yr[1], zr[0]:zr[1]]`` (0-based, half-open; default: the
        whole extent) packed on the device and copied to the host as float64 or float32
        (``ns3d_box_d2h``): ``Array(A)[2:end-1,2:end-1,2:end-1]`` is ``box(A, (1, sx-1), (1, sy-1), (1, sz-1))``."""
        sx, sy, sz = src.shape
        x0, x1 = xr if xr is not None else (0, sx)
        y0, y1 = yr if yr is not None else (0, sy)
        z0, z1 = zr if zr is not None else (0, sz)
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("box: dtype must be float64 or float32")
        out = np.empty((max(x1 - x0, 0), max(y1 - y0, 0), max(z1 - z0, 0)), dtype=dtype, order="F")
        self._ck(self.lib.ns3d_box_d2h(self.h, src.ptr, sx, sy, sz, x0, x1, y0, y1, z0, z1, out.ctypes.data,
                                       int(dtype == np.dtype(np.float32))), "ns3d_box_d2h")
        return out

    def gather_box(self, src: DeviceArray, xr, yr, zr, nplanes_all, dtype=np.float64):
        """``gather!`` of z-slabs (``ns3d_gather_box``): every rank passes its box, rank 0 gets the boxes
        concatenated along z (other ranks get None).  ``nplanes_all[r]`` = planes rank r contributes."""
        sx, sy, sz = src.shape
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("gather_box: dtype must be float64 or float32")
        rank = self.lib.ns3d_comm_rank(self.h)
        counts = (C.c_int * len(nplanes_all))(*nplanes_all)
        out = np.empty((xr[1] - xr[0], yr[1] - yr[0], int(sum(nplanes_all))), dtype=dtype, order="F") if rank == 0 else None
        self._ck(self.lib.ns3d_gather_box(self.h, src.ptr, sx, sy, sz, xr[0], xr[1], yr[0], yr[1], zr[0], zr[1], counts,
                                          out.ctypes.data if out is not None else None,
                                          int(dtype == np.dtype(np.float32))), "ns3d_gather_box")
        return out

    # -- asynchronous copies on the context's upload / download streams (pinned host memory) -----------
    def h2d_async(self, dst_ptr: int, host_ptr: int, count: int):
        self._ck(self.lib.ns3d_h2d_async(self.h, dst_ptr, host_ptr, count), "ns3d_h2d_async")

    def d2h_async(self, host_ptr: int, src_ptr: int, count: int):
        self._ck(self.lib.ns3d_d2h_async(self.h, host_ptr, src_ptr, count), "ns3d_d2h_async")

    def stream_wait(self, waiter: int, signaller: int):
        """Everything enqueued on stream ``waiter`` from now on waits for what ``signaller`` holds so far."""
        self._ck(self.lib.ns3d_stream_wait(self.h, waiter, signaller), "ns3d_stream_wait")

    def stream_sync(self, which: int):
        self._ck(self.lib.ns3d_stream_sync(self.h, which), "ns3d_stream_sync")

    def d2h_raw(self, host_ptr: int, src_ptr: int, count: int):
        self._ck(self.lib.ns3d_d2h(self.h, host_ptr, src_ptr, count), "ns3d_d2h")

    def from_host(self, host: np.ndarray) -> DeviceArray:
        return self.zeros(*host.shape).set(host)

    def fill_profile_z(self, a: DeviceArray, profile: np.ndarray):
        """``A[ix,iy,iz] = profile[iz]``: the scripts' z-dependent initial arrays (G:86-87, M:370), built on the device."""
        prof = np.ascontiguousarray(profile, dtype=np.float64)
        if prof.shape != (a.shape[2],):
            raise NS3DError(f"fill_profile_z: {prof.shape} values for {a.shape[2]} planes")
        self._ck(self.lib.ns3d_fill_profile_z(self.h, a.ptr, a.shape[0], a.shape[1], a.shape[2],
                                              prof.ctypes.data_as(c_double_p)), "ns3d_fill_profile_z")

    def fill_profile_zy(self, a: DeviceArray, profile: np.ndarray, add_y: np.ndarray, add_z: np.ndarray):
        """``A[ix,iy,iz] = (profile[iz] + add_y[iy]) + add_z[iz]`` (M:370 term by term: signed zeros when g = 0)."""
        prof, ay, az = (np.ascontiguousarray(v, dtype=np.float64) for v in (profile, add_y, add_z))
        if prof.shape != (a.shape[2],) or az.shape != (a.shape[2],) or ay.shape != (a.shape[1],):
            raise NS3DError(f"fill_profile_zy: {prof.shape}, {ay.shape}, {az.shape} values for an array of shape {a.shape}")
        self._ck(self.lib.ns3d_fill_profile_zy(self.h, a.ptr, a.shape[0], a.shape[1], a.shape[2], prof.ctypes.data_as(c_double_p),
                                               ay.ctypes.data_as(c_double_p), az.ctypes.data_as(c_double_p)), "ns3d_fill_profile_zy")

    def fill_plane_x(self, a: DeviceArray, ix: int, value: float):
        """``A[ix+1,:,:] .= value`` (M:369; ix 0-based)."""
        self._ck(self.lib.ns3d_fill_plane_x(self.h, a.ptr, a.shape[0], a.shape[1], a.shape[2], ix, value), "ns3d_fill_plane_x")

    def copy(self, dst: DeviceArray, src: DeviceArray):
        self.call("ns3d_copy", dst, src, src.size)

    def max_abs(self, a: DeviceArray) -> float:
        """``max_g(abs.(A))`` (M:21,466)."""
        out = C.c_double()
        self._ck(self.lib.ns3d_max_abs(self.h, a.ptr, a.size, C.byref(out)), "ns3d_max_abs")
        return out.value

    # -- communicator --------------------------------------------------------------------------
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.ns3d_comm_unique_id(buf)
        if rc != 0:
            raise NS3DError(f"ns3d_comm_unique_id failed ({rc}): {self.lib.ns3d_last_error(None).decode()}")
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        self._ck(self.lib.ns3d_comm_init(self.h, rank, nranks, uid), "ns3d_comm_init")

    def update_halo(self, fields, nz: int):
        """``update_halo!(A...)`` for z-slabs."""
        n = len(fields)
        ptrs = (C.c_void_p * n)(*[f.ptr for f in fields])
        sx = (C.c_int * n)(*[f.shape[0] for f in fields])
        sy = (C.c_int * n)(*[f.shape[1] for f in fields])
        sz = (C.c_int * n)(*[f.shape[2] for f in fields])
        self._ck(self.lib.ns3d_update_halo(self.h, ptrs, sx, sy, sz, n, nz), "ns3d_update_halo")

    def allreduce_max(self, x: float) -> float:
        v = C.c_double(x)
        self._ck(self.lib.ns3d_allreduce_max(self.h, C.byref(v)), "ns3d_allreduce_max")
        return v.value

    # -- level 2 -------------------------------------------------------------------------------
    def pt_solve(self, Pr, dPrdtau, divV, p: PtParams):
        """PT loop (M:458-471) -> (iterations, [err at each check])."""
        cap = max(p.niter // max(p.nchk, 1) + 2, 2)
        hist = (C.c_double * cap)()
        iters, nchecks = C.c_int(0), C.c_int(0)
        self._ck(self.lib.ns3d_pt_solve(self.h, _ptr(Pr), _ptr(dPrdtau), _ptr(divV), C.byref(p), C.byref(iters),
                                        hist, cap, C.byref(nchecks)), "ns3d_pt_solve")
        return iters.value, [hist[i] for i in range(min(nchecks.value, cap))]

    def pt_iterate(self, Pr, dPrdtau, divV, p: PtParams, n: int):
        self._ck(self.lib.ns3d_pt_iterate(self.h, _ptr(Pr), _ptr(dPrdtau), _ptr(divV), C.byref(p), n),
                 "ns3d_pt_iterate")

    def _describe(self, p: PtParams):
        buf, n = C.create_string_buffer(512), C.c_int(0)
        self._ck(self.lib.ns3d_pt_describe(self.h, C.byref(p), buf, 512, C.byref(n)), "ns3d_pt_describe")
        return buf.value.decode(), n.value

    def pt_kernel_name(self, p: PtParams) -> str:
        """The kernel the fused loop launches for ``p`` on this context (``ns3d_pt_describe``)."""
        return self._describe(p)[0]

    def pt_iters_per_launch(self, p: PtParams) -> int:
        return self._describe(p)[1]

    def predictor(self, fields: Fields, sp: StepParams):
        """M:449-455: update_τ!, predict_V!, set_cylinder!, update_∇V! + halo updates."""
        self._ck(self.lib.ns3d_predictor(self.h, C.byref(fields), C.byref(sp)), "ns3d_predictor")

    def corrector(self, fields: Fields, sp: StepParams):
        """M:472-474: correct_V!, set_cylinder!, set_bc_Vel!."""
        self._ck(self.lib.ns3d_corrector(self.h, C.byref(fields), C.byref(sp)), "ns3d_corrector")

    def advect_swap(self, fields: Fields, sp: StepParams):
        """M:475-477: the four snapshots, advect!, update_halo!(Vx,Vy,Vz)."""
        self._ck(self.lib.ns3d_advect_swap(self.h, C.byref(fields), C.byref(sp)), "ns3d_advect_swap")

    def step(self, fields: Fields, sp: StepParams):
        """One time step (M:449-477) -> (iterations, [err at each check])."""
        cap = max(sp.pt.niter // max(sp.pt.nchk, 1) + 2, 2)
        hist = (C.c_double * cap)()
        iters, nchecks = C.c_int(0), C.c_int(0)
        self._ck(self.lib.ns3d_step(self.h, C.byref(fields), C.byref(sp), C.byref(iters), hist, cap,
                                    C.byref(nchecks)), "ns3d_step")
        return iters.value, [hist[i] for i in range(min(nchecks.value, cap))]
