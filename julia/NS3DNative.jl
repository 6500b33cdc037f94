# NS3DNative.jl -- Julia `ccall` shim over libns3d.so (include/ns3d.h).
#
# Drop-in replacement of the ParallelStencil / ImplicitGlobalGrid / MPI call sites of
# scripts/NavierStokes3D_multi_gpu.jl (M) and scripts/NavierStokes3D_gpu.jl (G).  Two layers:
#
#   * explicit functions: every reference kernel under its own name and argument order with the
#     context prepended, e.g. `update_τ!(ctx, τxx, ..., dx, dy, dz)`;
#   * a ParallelStencil / ImplicitGlobalGrid look-alike surface on a default context, so that a run
#     script keeps its text: `@init_ns3d(device)`, `@zeros(nx,ny,nz)`, `@parallel update_τ!(...)`,
#     `@parallel (1:n, 1:m) bc_x!(A)`, `Data.Array(host)`, `Array(dev)`, `copy!(A_o, A)`,
#     `init_global_grid(nx,ny,nz)`, `nx_g()`, `x_g(ix,dx,A)`, `update_halo!(A...)`, `gather!(A_inn, A_v)`,
#     `max_g(abs, A)`, `finalize_global_grid()` -- see scripts/NavierStokes3D_b200.jl and
#     scripts/NavierStokes3D_gpu_b200.jl.
#
# NOT EXECUTED BY JULIA in the build container (Julia is not installed there).  What runs there instead:
# oracle/jl_shim.py interprets this file's text -- structs, typed methods, `Ref`/`Ptr`, every `ccall` converted by
# its declared Julia types and sent into libns3d.so -- together with the two run scripts; every ccall site below
# is reached and the scripts reproduce the reference's results bit for bit (tests/test_julia_shim_exec.py, CPU
# emulation and B200), and tests/test_julia_shim_static.py checks every ccall against include/ns3d.h.  The same
# C ABI is also exercised by the Python ctypes binding navierstokes3d_b200/native.py, which mirrors this file
# call for call.  Every `ccall` below passes its arguments one by one (`ccall` takes no splats).
module NS3DNative

export Ctx, DevArray, zeros3, to_host, set!, set_mode!, fill_profile_z!, fill_profile_zy!, fill_plane_x!, PARITY, FAST, FASTEST, VARIANT_M, VARIANT_G,
       update_τ!, predict_V!, update_∇V!, update_dPrdτ!, update_Pr!, compute_res!, max_g_abs, correct_V!,
       bc_x!, bc_y!, bc_z!, bc_x_Vx!, bc_x_Pr!, bc_zV!, bc_xhydstatic!, set_bc_Vel_M!, set_bc_Vel_G!,
       set_bc_Pr_M!, set_bc_Pr_G!, advect!, set_cylinder_M!, set_cylinder_G!, update_halo!,
       comm_init_mpi!, PtParams, pt_solve!, inner, inner32, plane_xy, plane_xz, gather_inner,
       Fields, StepParams, predictor!, corrector!, advect_swap!, step!,
       @init_ns3d, @zeros, @parallel, Data, default_ctx, init_global_grid, finalize_global_grid,
       nx_g, ny_g, nz_g, x_g, y_g, z_g, gather!, max_g

const LIB = get(ENV, "NS3D_LIB", joinpath(@__DIR__, "..", "navierstokes3d_b200", "csrc", "libns3d.so"))
const PARITY, FAST, FASTEST = Cint(0), Cint(1), Cint(2)
const VARIANT_M, VARIANT_G = Cint(0), Cint(1)

mutable struct Ctx
    h::Ptr{Cvoid}
end

struct DevArray            # a dense column-major Float64 device array of the reference's shape
    p::Ptr{Float64}
    dims::NTuple{3,Int}
end
Base.size(a::DevArray) = a.dims
Base.size(a::DevArray, d::Integer) = a.dims[d]
Base.length(a::DevArray) = prod(a.dims)

lasterr(h::Ptr{Cvoid}) = unsafe_string(ccall((:ns3d_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
function check(c::Ctx, rc)
    rc == 0 || error("libns3d: ", lasterr(c.h), " (", rc, ")")
    return nothing
end

"Replaces `@init_parallel_stencil(CUDA, Float64, 3)` (M:5) and IGG's GPU selection."
function Ctx(device::Integer=0; mode=FAST)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ns3d_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
    rc == 0 || error("ns3d_create: ", lasterr(Ptr{Cvoid}(C_NULL)))
    c = Ctx(r[])
    set_mode!(c, mode)
    finalizer(x -> ccall((:ns3d_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end
"Positional form: what `@init_ns3d(device, mode)` expands to."
Ctx(device::Integer, mode::Integer) = Ctx(device; mode=mode)
set_mode!(c::Ctx, m) = check(c, ccall((:ns3d_set_mode, LIB), Cint, (Ptr{Cvoid}, Cint), c.h, m))

"`@zeros(nx,ny,nz)` (M:343-360)"
function zeros3(c::Ctx, nx, ny, nz)
    r = Ref{Ptr{Float64}}(C_NULL)
    check(c, ccall((:ns3d_zeros, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ref{Ptr{Float64}}), c.h, nx, ny, nz, r))
    return DevArray(r[], (Int(nx), Int(ny), Int(nz)))
end
"`Data.Array(host)` (M:370)"
function set!(c::Ctx, a::DevArray, h::Array{Float64,3})
    size(h) == a.dims || error("shape mismatch")
    check(c, ccall((:ns3d_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Csize_t), c.h, a.p, h, length(h)))
    return a
end
"`Array(A)` (M:399)"
function to_host(c::Ctx, a::DevArray)
    h = Array{Float64,3}(undef, a.dims)
    check(c, ccall((:ns3d_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Csize_t), c.h, h, a.p, length(h)))
    return h
end
"`A_o .= A` (M:475): a method of `Base.copy!`, not a new generic"
function Base.copy!(c::Ctx, dst::DevArray, src::DevArray)
    dst.dims == src.dims || error("shape mismatch")
    check(c, ccall((:ns3d_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Csize_t), c.h, dst.p, src.p, length(src)))
    return dst
end

# ---- device-side initialisers: the comprehensions of G:86-87 / M:369-370 without a 3-D host array ----
"`A[ix,iy,iz] = profile[iz]` (G:86-87): nz host values cross PCIe"
function fill_profile_z!(c::Ctx, a::DevArray, profile::Vector{Float64})
    length(profile) == a.dims[3] || error("fill_profile_z!: one value per z-plane")
    sx, sy, sz = a.dims
    check(c, ccall((:ns3d_fill_profile_z, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Ptr{Float64}), c.h, a.p, sx, sy, sz, profile))
    return a
end
"`A[ix,iy,iz] = (profile[iz] + add_y[iy]) + add_z[iz]` (M:370 term by term: with g = 0 the terms are signed zeros)"
function fill_profile_zy!(c::Ctx, a::DevArray, profile::Vector{Float64}, add_y::Vector{Float64}, add_z::Vector{Float64})
    (length(profile) == a.dims[3] && length(add_z) == a.dims[3] && length(add_y) == a.dims[2]) || error("fill_profile_zy!: shape mismatch")
    sx, sy, sz = a.dims
    check(c, ccall((:ns3d_fill_profile_zy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, a.p, sx, sy, sz, profile, add_y, add_z))
    return a
end
"`A[ix,:,:] .= value` (M:369); ix 1-based"
function fill_plane_x!(c::Ctx, a::DevArray, ix::Integer, value::Real)
    sx, sy, sz = a.dims
    check(c, ccall((:ns3d_fill_plane_x, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cdouble), c.h, a.p, sx, sy, sz, ix - 1, value))
    return a
end

# ---- output path: only the requested box leaves the device (ns3d_box_d2h) ---------------------
function box(c::Ctx, a::DevArray, xr::UnitRange, yr::UnitRange, zr::UnitRange, ::Type{T}) where {T<:Union{Float64,Float32}}
    h = Array{T,3}(undef, length(xr), length(yr), length(zr))      # 1-based inclusive ranges -> 0-based half-open
    sx, sy, sz = a.dims
    check(c, ccall((:ns3d_box_d2h, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Cint),
                   c.h, a.p, sx, sy, sz, first(xr) - 1, last(xr), first(yr) - 1, last(yr), first(zr) - 1, last(zr), h,
                   T == Float32 ? 1 : 0))
    return h
end
"`Array(A)[2:end-1,2:end-1,2:end-1]` (M:399-403, 528-532)"
inner(c::Ctx, a::DevArray) = box(c, a, 2:a.dims[1]-1, 2:a.dims[2]-1, 2:a.dims[3]-1, Float64)
"`convert.(Float32, Array(A)[2:end-1,2:end-1,2:end-1])` (M:408): converted on the device"
inner32(c::Ctx, a::DevArray) = box(c, a, 2:a.dims[1]-1, 2:a.dims[2]-1, 2:a.dims[3]-1, Float32)
"`gather!(A_inn, A_v)` (M:399-403) on z-slabs: the global interior on rank 0 (`nothing` elsewhere).
 `nplanes[r+1]` = interior planes rank r contributes (nz-2; for `Vz` nz-1 on the last rank only)."
function gather_inner(c::Ctx, a::DevArray, me::Integer, nplanes::Vector{<:Integer}, ::Type{T}=Float64) where {T<:Union{Float64,Float32}}
    counts = Cint.(nplanes)
    sx, sy, sz = a.dims
    h = me == 0 ? Array{T,3}(undef, sx - 2, sy - 2, sum(nplanes)) : nothing
    hp = h === nothing ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(pointer(h))
    check(c, ccall((:ns3d_gather_box, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cvoid}, Cint),
                   c.h, a.p, sx, sy, sz, 1, sx - 1, 1, sy - 1, 1, 1 + nplanes[me+1], counts, hp, T == Float32 ? 1 : 0))
    return h
end
"`A_v[:, :, k]` / `A_v[:, j, :]` of the interior (heat-map planes, M:422-431); k, j index the interior"
plane_xy(c::Ctx, a::DevArray, k::Integer) = box(c, a, 2:a.dims[1]-1, 2:a.dims[2]-1, k+1:k+1, Float64)[:, :, 1]
plane_xz(c::Ctx, a::DevArray, j::Integer) = box(c, a, 2:a.dims[1]-1, j+1:j+1, 2:a.dims[3]-1, Float64)[:, 1, :]

const P = Ptr{Float64}

# ---- level 1: same names / argument order as the reference kernels -------------------------
# (nx, ny, nz) = size of the cell-centred fields; C cannot read size(A), so it is appended.
function update_τ!(c::Ctx, τxx, τyy, τzz, τxy, τxz, τyz, Vx, Vy, Vz, μ, dx, dy, dz)
    nx, ny, nz = τxx.dims
    check(c, ccall((:ns3d_update_tau, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, τxx.p, τyy.p, τzz.p, τxy.p, τxz.p, τyz.p, Vx.p, Vy.p, Vz.p, μ, dx, dy, dz, nx, ny, nz))
end
function predict_V!(c::Ctx, Vx, Vy, Vz, τxx, τyy, τzz, τxy, τxz, τyz, ρ, g, dt, dx, dy, dz)
    nx, ny, nz = τxx.dims
    check(c, ccall((:ns3d_predict_V, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, τxx.p, τyy.p, τzz.p, τxy.p, τxz.p, τyz.p, ρ, g, dt, dx, dy, dz, nx, ny, nz))
end
function update_∇V!(c::Ctx, ∇V, Vx, Vy, Vz, dx, dy, dz)
    nx, ny, nz = ∇V.dims
    check(c, ccall((:ns3d_update_divV, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, ∇V.p, Vx.p, Vy.p, Vz.p, dx, dy, dz, nx, ny, nz))
end
function update_dPrdτ!(c::Ctx, Pr, dPrdτ, ∇V, ρ, dt, dτ, damp, dx, dy, dz)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_update_dPrdtau, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Pr.p, dPrdτ.p, ∇V.p, ρ, dt, dτ, damp, dx, dy, dz, nx, ny, nz))
end
function update_Pr!(c::Ctx, Pr, dPrdτ, dτ)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_update_Pr, LIB), Cint, (Ptr{Cvoid}, P, P, Cdouble, Cint, Cint, Cint), c.h, Pr.p, dPrdτ.p, dτ, nx, ny, nz))
end
function compute_res!(c::Ctx, Rp, Pr, ∇V, ρ, dt, dx, dy, dz)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_compute_res, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Rp.p, Pr.p, ∇V.p, ρ, dt, dx, dy, dz, nx, ny, nz))
end
"`max_g(abs.(A))` (M:21,466) / `maximum(abs.(A))` (G:132): device reduction + NCCL max-allreduce, NaN-propagating"
function max_g_abs(c::Ctx, A::DevArray)
    r = Ref{Cdouble}(0.0)
    check(c, ccall((:ns3d_max_abs, LIB), Cint, (Ptr{Cvoid}, P, Csize_t, Ref{Cdouble}), c.h, A.p, length(A), r))
    return r[]
end
function correct_V!(c::Ctx, Vx, Vy, Vz, Pr, dt, ρ, dx, dy, dz)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_correct_V, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, Pr.p, dt, ρ, dx, dy, dz, nx, ny, nz))
end
# face kernels: the array's OWN shape goes to C (bc_x!(Vx) works on an (nx+1,ny,nz) array)
function bc_x!(c::Ctx, A)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_x, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cint, Cint), c.h, A.p, sx, sy, sz))
end
function bc_y!(c::Ctx, A)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_y, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cint, Cint), c.h, A.p, sx, sy, sz))
end
function bc_z!(c::Ctx, A)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_z, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cint, Cint), c.h, A.p, sx, sy, sz))
end
function bc_zV!(c::Ctx, A)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_zV, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cint, Cint), c.h, A.p, sx, sy, sz))
end
function bc_x_Vx!(c::Ctx, A, V)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_x_Vx, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cint, Cint), c.h, A.p, V, sx, sy, sz))
end
function bc_x_Pr!(c::Ctx, A, val)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_x_Pr, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cint, Cint), c.h, A.p, val, sx, sy, sz))
end
function bc_xhydstatic!(c::Ctx, A, dz, nz, g, ρ)
    sx, sy, sz = A.dims
    check(c, ccall((:ns3d_bc_xhydstatic, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, A.p, dz, nz, g, ρ, sx, sy, sz))
end
"`set_bc_Vel!(Vx,Vy,Vz,xvo_g,lx,vin)` (M:156): the float == guard of M:164 is evaluated HERE, as written"
function set_bc_Vel_M!(c::Ctx, Vx, Vy, Vz, xvo_g, lx, vin, Pr)
    nx, ny, nz = Pr.dims
    guard = xvo_g == -lx / 2 ? 1 : 0
    check(c, ccall((:ns3d_set_bc_Vel_M, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cint, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, guard, vin, nx, ny, nz))
end
function set_bc_Vel_G!(c::Ctx, Vx, Vy, Vz, Pr)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_set_bc_Vel_G, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cint, Cint, Cint), c.h, Vx.p, Vy.p, Vz.p, nx, ny, nz))
end
"`set_bc_Pr!(Pr, xve_g, lx, val)` (M:175): guard of M:179 evaluated here"
function set_bc_Pr_M!(c::Ctx, Pr, xve_g, lx, val)
    nx, ny, nz = Pr.dims
    guard = xve_g == lx / 2 ? 1 : 0
    check(c, ccall((:ns3d_set_bc_Pr_M, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cdouble, Cint, Cint, Cint), c.h, Pr.p, guard, val, nx, ny, nz))
end
"`set_bc_Pr!(Pr, dz, nz, g, ρ)` (G:281)"
function set_bc_Pr_G!(c::Ctx, Pr, dz, nz_arg, g, ρ)
    nx, ny, nz = Pr.dims
    check(c, ccall((:ns3d_set_bc_Pr_G, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Pr.p, dz, nz_arg, g, ρ, nx, ny, nz))
end
function advect!(c::Ctx, Vx, Vx_o, Vy, Vy_o, Vz, Vz_o, C, C_o, dt, dx, dy, dz)
    nx, ny, nz = C.dims
    check(c, ccall((:ns3d_advect, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vx_o.p, Vy.p, Vy_o.p, Vz.p, Vz_o.p, C.p, C_o.p, dt, dx, dy, dz, nx, ny, nz))
end
"`set_cylinder!` of script M (M:249); zco_g, lx, ly, lz, dz are dead arguments there and are dropped"
function set_cylinder_M!(c::Ctx, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, zco_g, lx, ly, lz, dx, dy, dz)
    nx, ny, nz = C.dims
    check(c, ccall((:ns3d_set_cylinder_M, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, C.p, Vx.p, Vy.p, Vz.p, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, dx, dy, nx, ny, nz))
end
"`set_cylinder!` of script G (G:336)"
function set_cylinder_G!(c::Ctx, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, lx, ly, lz, dx, dy, dz)
    nx, ny, nz = C.dims
    check(c, ccall((:ns3d_set_cylinder_G, LIB), Cint,
                   (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, C.p, Vx.p, Vy.p, Vz.p, a2, b2, ox, oy, sinβ, cosβ, lx, ly, dx, dy, nx, ny, nz))
end

# ---- communication: replaces ImplicitGlobalGrid + MPI.Allreduce ------------------------------
"Attach the library's NCCL communicator to an MPI.jl job (dims = (1,1,nprocs), z-slabs).
 Rank 0 creates the 128-byte id, MPI broadcasts it -- the only use of MPI that remains."
function comm_init_mpi!(c::Ctx, MPI, comm)
    id = zeros(UInt8, 128)
    me, np = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    if me == 0
        rc = ccall((:ns3d_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)
        rc == 0 || error("ns3d_comm_unique_id failed (", rc, ")")
    end
    MPI.Bcast!(id, 0, comm)
    check(c, ccall((:ns3d_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), c.h, me, np, id))
    return me, (1, 1, np)
end
"`update_halo!(A...)`; nz = local cell count along the split dimension"
function update_halo!(c::Ctx, nz::Integer, A::DevArray...)
    ps = P[a.p for a in A]
    sx = Cint[a.dims[1] for a in A]
    sy = Cint[a.dims[2] for a in A]
    sz = Cint[a.dims[3] for a in A]
    check(c, ccall((:ns3d_update_halo, LIB), Cint, (Ptr{Cvoid}, Ptr{P}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Cint, Cint),
                   c.h, ps, sx, sy, sz, length(A), nz))
end

# ---- level 2: the fused pseudo-transient loop (M:458-471 / G:126-137) -------------------------
struct PtParams          # ns3d_pt_params, field for field
    nx::Cint; ny::Cint; nz::Cint; variant::Cint
    rho::Cdouble; dt::Cdouble; dtau::Cdouble; damp::Cdouble; dx::Cdouble; dy::Cdouble; dz::Cdouble
    eps_it::Cdouble; err_num::Cdouble; err_den::Cdouble
    niter::Cint; nchk::Cint; outlet_guard::Cint
    outlet_val::Cdouble; g::Cdouble
    zchunk::Cint; reserved::Cint
end
"Returns (iterations, err history) -- what the script pushes into `err_evo` (M:467)."
function pt_solve!(c::Ctx, Pr, dPrdτ, ∇V, p::PtParams)
    cap = p.niter ÷ max(p.nchk, 1) + 2
    hist = zeros(Cdouble, cap)
    it = Ref{Cint}(0)
    nc = Ref{Cint}(0)
    check(c, ccall((:ns3d_pt_solve, LIB), Cint, (Ptr{Cvoid}, P, P, P, Ref{PtParams}, Ref{Cint}, Ptr{Cdouble}, Cint, Ref{Cint}),
                   c.h, Pr.p, dPrdτ.p, ∇V.p, p, it, hist, cap, nc))
    return Int(it[]), hist[1:nc[]]
end

# ---- level 2: the once-per-step groups and the whole step (M:449-477 / G:121-142) ------------------
struct Fields            # ns3d_fields: the 18 arrays in the script's allocation order (M:343-360)
    Pr::P; dPrdtau::P; C::P; C_o::P; txx::P; tyy::P; tzz::P; txy::P; txz::P; tyz::P
    Vx::P; Vy::P; Vz::P; Vx_o::P; Vy_o::P; Vz_o::P; divV::P; Rp::P
end
function Fields(Pr::DevArray, dPrdτ::DevArray, C::DevArray, C_o::DevArray, τxx::DevArray, τyy::DevArray, τzz::DevArray,
                τxy::DevArray, τxz::DevArray, τyz::DevArray, Vx::DevArray, Vy::DevArray, Vz::DevArray, Vx_o::DevArray,
                Vy_o::DevArray, Vz_o::DevArray, ∇V::DevArray, Rp::DevArray)
    return Fields(Pr.p, dPrdτ.p, C.p, C_o.p, τxx.p, τyy.p, τzz.p, τxy.p, τxz.p, τyz.p, Vx.p, Vy.p, Vz.p, Vx_o.p, Vy_o.p,
                  Vz_o.p, ∇V.p, Rp.p)
end
struct StepParams        # ns3d_step_params, field for field
    pt::PtParams
    mu::Cdouble; vin::Cdouble
    a2::Cdouble; b2::Cdouble; ox::Cdouble; oy::Cdouble; sinb::Cdouble; cosb::Cdouble
    xco_g::Cdouble; yco_g::Cdouble; lx::Cdouble; ly::Cdouble
    inlet_guard::Cint; reserved::Cint
end
predictor!(c::Ctx, f::Fields, sp::StepParams) =
    check(c, ccall((:ns3d_predictor, LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}), c.h, f, sp))
corrector!(c::Ctx, f::Fields, sp::StepParams) =
    check(c, ccall((:ns3d_corrector, LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}), c.h, f, sp))
advect_swap!(c::Ctx, f::Fields, sp::StepParams) =
    check(c, ccall((:ns3d_advect_swap, LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}), c.h, f, sp))
"One time step = predictor!, pt_solve!, corrector!, advect_swap!; returns (iterations, err history)."
function step!(c::Ctx, f::Fields, sp::StepParams)
    cap = sp.pt.niter ÷ max(sp.pt.nchk, 1) + 2
    hist = zeros(Cdouble, cap)
    it = Ref{Cint}(0)
    nc = Ref{Cint}(0)
    check(c, ccall((:ns3d_step, LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}, Ref{Cint}, Ptr{Cdouble}, Cint, Ref{Cint}),
                   c.h, f, sp, it, hist, cap, nc))
    return Int(it[]), hist[1:nc[]]
end

# =================================================================================================
# ParallelStencil / ImplicitGlobalGrid look-alike surface on a default context
# =================================================================================================
const DEFAULT = Ref{Union{Nothing,Ctx}}(nothing)
"The context the macros below act on (created by `@init_ns3d` / `init_global_grid`)."
function default_ctx()
    DEFAULT[] === nothing && error("NS3DNative: call @init_ns3d(device) or init_global_grid(nx,ny,nz) first")
    return DEFAULT[]::Ctx
end

"`@init_ns3d(device, mode)` -- stands where `@init_parallel_stencil(CUDA, Float64, 3)` stood (M:5, G:5)."
macro init_ns3d(args...)
    return :(NS3DNative.DEFAULT[] = NS3DNative.Ctx($(map(esc, args)...)))
end

"`@zeros(nx,ny,nz)` (M:343-360): a zero-filled device array of the library's allocator."
macro zeros(args...)
    return :(NS3DNative.zeros3(NS3DNative.default_ctx(), $(map(esc, args)...)))
end

"`@parallel f!(args...)` and `@parallel (ranges...) f!(args...)` (M:157-165, 449-476): the call itself, with
 the default context in front; ParallelStencil's launch ranges are implied by the arrays' shapes here."
macro parallel(args...)
    call = args[end]
    (call isa Expr && call.head == :call) || error("@parallel expects a function call")
    f = esc(call.args[1])
    rest = map(esc, call.args[2:end])
    return :($f(NS3DNative.default_ctx(), $(rest...)))
end

"`Data.Array(host)` (M:370, G:85-88) and `Data.Number`"
module Data
import ..NS3DNative
const Number = Float64
function Array(h::Base.Array{Float64,3})
    c = NS3DNative.default_ctx()
    a = NS3DNative.zeros3(c, size(h, 1), size(h, 2), size(h, 3))
    return NS3DNative.set!(c, a, h)
end
end # module Data

"`Array(dev)` (M:399, G:89): device -> host"
Base.Array(a::DevArray) = to_host(default_ctx(), a)
"`copy!(A_o, A)` for `A_o .= A` (M:475, G:141)"
Base.copy!(dst::DevArray, src::DevArray) = copy!(default_ctx(), dst, src)

# ---- ImplicitGlobalGrid look-alike for dims = (1, 1, nprocs) (z-slabs, overlap 2, halo width 1) ----
mutable struct Grid
    nxyz::NTuple{3,Int}
    dims::NTuple{3,Int}
    coords::NTuple{3,Int}
    me::Int
    nprocs::Int
    mpi::Any          # the MPI module the job was started with (nothing on a single rank)
end
const GRID = Ref{Union{Nothing,Grid}}(nothing)
grid() = GRID[] === nothing ? error("NS3DNative: init_global_grid(nx,ny,nz) has not been called") : GRID[]::Grid

"`me, dims, nprocs, coords, comm = init_global_grid(nx, ny, nz)` (M:325).  With `MPI` (the MPI.jl module, initialised by
 the caller or here) the job's ranks become z-slabs, one per GPU of the node; without it a single rank.
 `device` defaults to the node-local rank, like IGG's GPU selection."
function init_global_grid(nx::Integer, ny::Integer, nz::Integer; MPI=nothing, mode=FAST, device=nothing, quiet=false)
    me, np, comm = 0, 1, nothing
    if MPI !== nothing
        MPI.Initialized() || MPI.Init()
        comm = MPI.COMM_WORLD
        me, np = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    end
    dev = device === nothing ? me : device          # one rank per GPU of ONE node
    c = Ctx(dev; mode=mode)
    DEFAULT[] = c
    if MPI !== nothing
        comm_init_mpi!(c, MPI, comm)
    end
    GRID[] = Grid((Int(nx), Int(ny), Int(nz)), (1, 1, np), (0, 0, me), me, np, MPI)
    if me == 0 && !quiet
        println("Global grid: ", nx_g(), "x", ny_g(), "x", nz_g(), " (nprocs: ", np, ", dims: 1x1x", np, ")")
    end
    return me, (1, 1, np), np, (0, 0, me), comm
end
"`finalize_global_grid()` (M:534)"
function finalize_global_grid(; finalize_MPI=true)
    g = grid()
    DEFAULT[] = nothing
    GRID[] = nothing
    if g.mpi !== nothing && finalize_MPI
        g.mpi.Finalize()
    end
    return nothing
end
"`nx_g()`, `ny_g()`, `nz_g()` (M:328,338): dims*(n - overlap) + overlap"
nx_g() = (g = grid(); g.dims[1] * (g.nxyz[1] - 2) + 2)
ny_g() = (g = grid(); g.dims[2] * (g.nxyz[2] - 2) + 2)
nz_g() = (g = grid(); g.dims[3] * (g.nxyz[3] - 2) + 2)
function _x_g(i::Integer, d, A::DevArray, dim::Int)
    g = grid()
    n = g.nxyz[dim]
    x0 = 0.5 * (n - size(A, dim)) * d
    return (g.coords[dim] * (n - 2) + (i - 1)) * d + x0
end
"`x_g(ix,dx,A)`, `y_g`, `z_g` (M:363-367)"
x_g(i::Integer, d, A::DevArray) = _x_g(i, d, A, 1)
y_g(i::Integer, d, A::DevArray) = _x_g(i, d, A, 2)
z_g(i::Integer, d, A::DevArray) = _x_g(i, d, A, 3)
"`update_halo!(A...)` (10 call sites, SURVEY.md 2.2) on the default context"
update_halo!(A::DevArray...) = update_halo!(default_ctx(), grid().nxyz[3], A...)
"`max_g(abs.(A))` (M:21,466) is written `max_g(abs, A)` here: one device pass, no temporary, NCCL max-allreduce"
max_g(::typeof(abs), A::DevArray) = max_g_abs(default_ctx(), A)
"`gather!(A_inn, A_v)` (M:399-403): `A_inn` is not needed (the interior is extracted on the device); pass the field itself.
 Returns the gathered interior on rank 0, `nothing` elsewhere; `A_v`, when given on rank 0, is filled."
function gather!(A::DevArray, A_v=nothing; T=Float64)
    g = grid()
    c = default_ctx()
    if g.nprocs == 1
        out = T == Float32 ? inner32(c, A) : inner(c, A)
    else
        stag = size(A, 3) - g.nxyz[3]                       # 1 for Vz: the last rank contributes one more plane
        nplanes = [size(A, 3) - 2 - ((stag == 1 && r < g.nprocs - 1) ? 1 : 0) for r in 0:g.nprocs-1]
        out = gather_inner(c, A, g.me, nplanes, T)
    end
    if out !== nothing && A_v !== nothing
        A_v .= out
    end
    return out
end

end # module
