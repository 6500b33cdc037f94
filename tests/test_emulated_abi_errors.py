"""Error behaviour of the C ABI, exercised on the CPU through the emulated library (tests/emu):
every entry point returns a negative NS3D_E* code with a message in ns3d_last_error and never
aborts; the Python binding turns that into NS3DError -- the way a failing `@parallel` launch
surfaces as a Julia exception in the reference."""
import ctypes as C

import numpy as np
import pytest

from tests import emu


@pytest.fixture(scope="module", autouse=True)
def _emulated_library():
    with emu.use_emulated_library():
        yield


def test_unknown_option_and_bad_values(ns, ctx):
    with pytest.raises(ns.NS3DError, match="unknown option"):
        ctx.set_option("no_such_knob", 1)
    for name, bad in (("ptv_k", 4), ("ptv_ns", 2), ("ptv_lb", 2), ("ptv_lb", 9)):
        with pytest.raises(ns.NS3DError, match=name):
            ctx.set_option(name, bad)
    for name, ok in (("ptv_k", 3), ("ptv_ns", 5), ("ptv_lb", 4), ("ptv_tma", 0), ("ptv_pxt", 8), ("ptv_bty", 0), ("p2p_halo", 0),
                     ("graphs", 0), ("serpentine", -1)):
        ctx.set_option(name, ok)


def test_bad_mode(ns, ctx):
    with pytest.raises(ns.NS3DError, match="unknown mode"):
        ctx.set_mode(7)
    assert ctx.lib.ns3d_get_mode(ctx.h) == ns.PARITY


def test_zeros_rejects_bad_shapes_and_free_rejects_foreign_pointers(ns, ctx):
    with pytest.raises(ns.NS3DError, match="bad shape"):
        ctx.zeros(0, 4, 4)
    a = ctx.zeros(4, 4, 4)
    assert ctx.lib.ns3d_free(ctx.h, C.c_void_p(a.ptr + 8)) == -1          # NS3D_EINVAL
    assert b"not owned" in ctx.lib.ns3d_last_error(ctx.h)
    before = ctx.lib.ns3d_bytes_allocated(ctx.h)
    ctx.free(a)
    assert ctx.lib.ns3d_bytes_allocated(ctx.h) < before


def test_fused_loop_rejects_bad_arguments(ns, ctx):
    s = ns.setup_multi_gpu(12, ny=9, nz=9)
    Pr, dP, dv = ctx.zeros(12, 9, 9), ctx.zeros(10, 7, 7), ctx.zeros(12, 9, 9)
    pt = s.pt_params()
    pt.nchk = 0
    with pytest.raises(ns.NS3DError, match="niter/nchk"):
        ctx.pt_solve(Pr, dP, dv, pt)
    pt = s.pt_params()
    pt.variant = 5
    with pytest.raises(ns.NS3DError, match="unknown variant"):
        ctx.pt_iterate(Pr, dP, dv, pt, 2)
    pt = s.pt_params()
    pt.nx = 2
    with pytest.raises(ns.NS3DError, match="at least 3"):
        ctx.pt_iterate(Pr, dP, dv, pt, 2)
    # the loop runs on pitched copies: the caller's arrays are only read and written within their bounds (pack /
    # unpack), so any device array of the right shape will do; NULL is refused
    pt = s.pt_params()
    assert ctx.lib.ns3d_pt_iterate(ctx.h, None, dP.ptr, dv.ptr, C.byref(pt), 2) == -1


def test_null_context_and_null_arguments(ns, ctx):
    lib = ctx.lib
    assert lib.ns3d_sync(None) == -1 and lib.ns3d_set_mode(None, 0) == -1
    assert lib.ns3d_step(ctx.h, None, None, None, None, 0, None) == -1
    assert lib.ns3d_predictor(ctx.h, None, None) == -1 and lib.ns3d_corrector(ctx.h, None, None) == -1
    assert lib.ns3d_advect_swap(ctx.h, None, None) == -1
    assert lib.ns3d_launch_count(None) == 0 and lib.ns3d_bytes_allocated(None) == 0
    with pytest.raises(ns.NS3DError, match="bad array"):
        ctx.call("ns3d_bc_x", ctx.zeros(2, 4, 4), 2, 4, 4)


def test_max_abs_propagates_nan_like_julia(ns, ctx):
    a = np.asfortranarray(np.linspace(-3, 2, 60).reshape(5, 4, 3))
    d = ctx.from_host(a)
    assert ctx.max_abs(d) == 3.0
    a[2, 1, 1] = np.nan
    d.set(a)
    assert np.isnan(ctx.max_abs(d))                     # maximum(abs.(A)) propagates NaN (M:466,469)
    a[2, 1, 1] = -np.inf
    d.set(a)
    assert ctx.max_abs(d) == np.inf
