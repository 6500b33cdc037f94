# NS3DNative.jl -- Julia `ccall` shim over libns3d.so (include/ns3d.h).
#
# Drop-in replacement of the ParallelStencil / ImplicitGlobalGrid / MPI call sites of
# scripts/NavierStokes3D_multi_gpu.jl (M) and scripts/NavierStokes3D_gpu.jl (G): every function
# below has the name and the argument order of the reference kernel it replaces, so a driver
# keeps its structure and only drops the `@parallel` prefix (see scripts/NavierStokes3D_b200.jl).
#
# NOT EXECUTED in the build container (Julia is not installed there); the same C ABI is
# exercised end to end by the Python ctypes binding navierstokes3d_b200/native.py, which mirrors
# this file call for call.
module NS3DNative

export Ctx, DevArray, zeros3, to_host, set!, set_mode!, PARITY, FAST, FASTEST,
       update_τ!, predict_V!, update_∇V!, update_dPrdτ!, update_Pr!, compute_res!, max_g_abs, correct_V!,
       bc_x!, bc_y!, bc_z!, bc_x_Vx!, bc_x_Pr!, bc_zV!, bc_xhydstatic!, set_bc_Vel_M!, set_bc_Vel_G!,
       set_bc_Pr_M!, set_bc_Pr_G!, advect!, set_cylinder_M!, set_cylinder_G!, update_halo!, copy!,
       comm_init_mpi!, PtParams, pt_solve!, inner, inner32, plane_xy, plane_xz, gather_inner,
       Fields, StepParams, predictor!, corrector!, advect_swap!, step!

const LIB = get(ENV, "NS3D_LIB", joinpath(@__DIR__, "..", "navierstokes3d_b200", "csrc", "libns3d.so"))
const PARITY, FAST, FASTEST = Cint(0), Cint(1), Cint(2)

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

lasterr(c) = unsafe_string(ccall((:ns3d_last_error, LIB), Cstring, (Ptr{Cvoid},), c))
check(c::Ctx, rc) = rc == 0 || error("libns3d: ", lasterr(c.h), " (", rc, ")")

"Replaces `@init_parallel_stencil(CUDA, Float64, 3)` (M:5) and IGG's GPU selection."
function Ctx(device::Integer=0; mode=FAST)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ns3d_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
    rc == 0 || error("ns3d_create: ", lasterr(C_NULL))
    c = Ctx(r[])
    set_mode!(c, mode)
    finalizer(x -> ccall((:ns3d_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end
set_mode!(c::Ctx, m) = check(c, ccall((:ns3d_set_mode, LIB), Cint, (Ptr{Cvoid}, Cint), c.h, m))

"`@zeros(nx,ny,nz)` (M:343-360)"
function zeros3(c::Ctx, nx, ny, nz)
    r = Ref{Ptr{Float64}}(C_NULL)
    check(c, ccall((:ns3d_zeros, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ref{Ptr{Float64}}), c.h, nx, ny, nz, r))
    return DevArray(r[], (nx, ny, nz))
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
"`A_o .= A` (M:475)"
copy!(c::Ctx, dst::DevArray, src::DevArray) =
    check(c, ccall((:ns3d_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Csize_t), c.h, dst.p, src.p, length(src)))

# ---- output path: only the requested box leaves the device (ns3d_box_d2h) ---------------------
function box(c::Ctx, a::DevArray, xr::UnitRange, yr::UnitRange, zr::UnitRange, ::Type{T}) where {T<:Union{Float64,Float32}}
    h = Array{T,3}(undef, length(xr), length(yr), length(zr))      # 1-based inclusive ranges -> 0-based half-open
    check(c, ccall((:ns3d_box_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Cint),
                   c.h, a.p, a.dims..., first(xr) - 1, last(xr), first(yr) - 1, last(yr), first(zr) - 1, last(zr), h, T == Float32))
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
    h = me == 0 ? Array{T,3}(undef, a.dims[1] - 2, a.dims[2] - 2, sum(nplanes)) : nothing
    check(c, ccall((:ns3d_gather_box, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cvoid}, Cint),
                   c.h, a.p, a.dims..., 1, a.dims[1] - 1, 1, a.dims[2] - 1, 1, 1 + nplanes[me+1], counts,
                   h === nothing ? C_NULL : pointer(h), T == Float32))
    return h
end
"`A_v[:, :, k]` / `A_v[:, j, :]` of the interior (heat-map planes, M:422-431); k, j index the interior"
plane_xy(c::Ctx, a::DevArray, k::Integer) = box(c, a, 2:a.dims[1]-1, 2:a.dims[2]-1, k+1:k+1, Float64)[:, :, 1]
plane_xz(c::Ctx, a::DevArray, j::Integer) = box(c, a, 2:a.dims[1]-1, j+1:j+1, 2:a.dims[3]-1, Float64)[:, 1, :]

const P = Ptr{Float64}
n3(Pr::DevArray) = (Cint(Pr.dims[1]), Cint(Pr.dims[2]), Cint(Pr.dims[3]))

# ---- level 1: same names / argument order as the reference kernels -------------------------
update_τ!(c, τxx, τyy, τzz, τxy, τxz, τyz, Vx, Vy, Vz, μ, dx, dy, dz) = (n = n3(τxx);
    check(c, ccall((:ns3d_update_tau, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, τxx.p, τyy.p, τzz.p, τxy.p, τxz.p, τyz.p, Vx.p, Vy.p, Vz.p, μ, dx, dy, dz, n...)))
predict_V!(c, Vx, Vy, Vz, τxx, τyy, τzz, τxy, τxz, τyz, ρ, g, dt, dx, dy, dz) = (n = n3(τxx);
    check(c, ccall((:ns3d_predict_V, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, τxx.p, τyy.p, τzz.p, τxy.p, τxz.p, τyz.p, ρ, g, dt, dx, dy, dz, n...)))
update_∇V!(c, ∇V, Vx, Vy, Vz, dx, dy, dz) = (n = n3(∇V);
    check(c, ccall((:ns3d_update_divV, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, ∇V.p, Vx.p, Vy.p, Vz.p, dx, dy, dz, n...)))
update_dPrdτ!(c, Pr, dPrdτ, ∇V, ρ, dt, dτ, damp, dx, dy, dz) = (n = n3(Pr);
    check(c, ccall((:ns3d_update_dPrdtau, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Pr.p, dPrdτ.p, ∇V.p, ρ, dt, dτ, damp, dx, dy, dz, n...)))
update_Pr!(c, Pr, dPrdτ, dτ) = (n = n3(Pr);
    check(c, ccall((:ns3d_update_Pr, LIB), Cint, (Ptr{Cvoid}, P, P, Cdouble, Cint, Cint, Cint), c.h, Pr.p, dPrdτ.p, dτ, n...)))
compute_res!(c, Rp, Pr, ∇V, ρ, dt, dx, dy, dz) = (n = n3(Pr);
    check(c, ccall((:ns3d_compute_res, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Rp.p, Pr.p, ∇V.p, ρ, dt, dx, dy, dz, n...)))
"`max_g(abs.(A))` (M:21,466): device reduction + NCCL max-allreduce, NaN-propagating"
function max_g_abs(c, A::DevArray)
    r = Ref{Cdouble}(0.0)
    check(c, ccall((:ns3d_max_abs, LIB), Cint, (Ptr{Cvoid}, P, Csize_t, Ref{Cdouble}), c.h, A.p, length(A), r))
    return r[]
end
correct_V!(c, Vx, Vy, Vz, Pr, dt, ρ, dx, dy, dz) = (n = n3(Pr);
    check(c, ccall((:ns3d_correct_V, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, Pr.p, dt, ρ, dx, dy, dz, n...)))
for (jl, sym) in ((:bc_x!, :ns3d_bc_x), (:bc_y!, :ns3d_bc_y), (:bc_z!, :ns3d_bc_z), (:bc_zV!, :ns3d_bc_zV))
    @eval $jl(c, A) = check(c, ccall(($(QuoteNode(sym)), LIB), Cint, (Ptr{Cvoid}, P, Cint, Cint, Cint), c.h, A.p, n3(A)...))
end
bc_x_Vx!(c, A, V) = check(c, ccall((:ns3d_bc_x_Vx, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cint, Cint), c.h, A.p, V, n3(A)...))
bc_x_Pr!(c, A, val) = check(c, ccall((:ns3d_bc_x_Pr, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cint, Cint), c.h, A.p, val, n3(A)...))
bc_xhydstatic!(c, A, dz, nz, g, ρ) =
    check(c, ccall((:ns3d_bc_xhydstatic, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Cint), c.h, A.p, dz, nz, g, ρ, n3(A)...))
"`set_bc_Vel!(Vx,Vy,Vz,xvo_g,lx,vin)` (M:156): the float == guard of M:164 is evaluated HERE, as written"
set_bc_Vel_M!(c, Vx, Vy, Vz, xvo_g, lx, vin, Pr) =
    check(c, ccall((:ns3d_set_bc_Vel_M, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cint, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vy.p, Vz.p, xvo_g == -lx / 2, vin, n3(Pr)...))
set_bc_Vel_G!(c, Vx, Vy, Vz, Pr) =
    check(c, ccall((:ns3d_set_bc_Vel_G, LIB), Cint, (Ptr{Cvoid}, P, P, P, Cint, Cint, Cint), c.h, Vx.p, Vy.p, Vz.p, n3(Pr)...))
"`set_bc_Pr!(Pr, xve_g, lx, val)` (M:175): guard of M:179 evaluated here"
set_bc_Pr_M!(c, Pr, xve_g, lx, val) =
    check(c, ccall((:ns3d_set_bc_Pr_M, LIB), Cint, (Ptr{Cvoid}, P, Cint, Cdouble, Cint, Cint, Cint), c.h, Pr.p, xve_g == lx / 2, val, n3(Pr)...))
set_bc_Pr_G!(c, Pr, dz, nz, g, ρ) =
    check(c, ccall((:ns3d_set_bc_Pr_G, LIB), Cint, (Ptr{Cvoid}, P, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Cint), c.h, Pr.p, dz, nz, g, ρ, n3(Pr)...))
advect!(c, Vx, Vx_o, Vy, Vy_o, Vz, Vz_o, C, C_o, dt, dx, dy, dz) = (n = n3(C);
    check(c, ccall((:ns3d_advect, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, Vx.p, Vx_o.p, Vy.p, Vy_o.p, Vz.p, Vz_o.p, C.p, C_o.p, dt, dx, dy, dz, n...)))
"`set_cylinder!` of script M (M:249); zco_g, lx, ly, lz, dz are dead arguments there and are dropped"
set_cylinder_M!(c, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, zco_g, lx, ly, lz, dx, dy, dz) = (n = n3(C);
    check(c, ccall((:ns3d_set_cylinder_M, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, C.p, Vx.p, Vy.p, Vz.p, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, dx, dy, n...)))
"`set_cylinder!` of script G (G:336)"
set_cylinder_G!(c, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, lx, ly, lz, dx, dy, dz) = (n = n3(C);
    check(c, ccall((:ns3d_set_cylinder_G, LIB), Cint, (Ptr{Cvoid}, P, P, P, P, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint, Cint, Cint),
                   c.h, C.p, Vx.p, Vy.p, Vz.p, a2, b2, ox, oy, sinβ, cosβ, lx, ly, dx, dy, n...)))

# ---- communication: replaces ImplicitGlobalGrid + MPI.Allreduce ------------------------------
"Attach the library's NCCL communicator to an MPI.jl job (dims = (1,1,nprocs), z-slabs).
 Rank 0 creates the 128-byte id, MPI broadcasts it -- the only use of MPI that remains."
function comm_init_mpi!(c::Ctx, MPI, comm)
    id = zeros(UInt8, 128)
    me, np = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    me == 0 && ccall((:ns3d_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || me != 0 || error("ns3d_comm_unique_id")
    MPI.Bcast!(id, 0, comm)
    check(c, ccall((:ns3d_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), c.h, me, np, id))
    return me, (1, 1, np)
end
"`update_halo!(A...)`; nz = local cell count along the split dimension"
function update_halo!(c::Ctx, nz::Integer, A::DevArray...)
    ps = P[a.p for a in A]
    sx = Cint[a.dims[1] for a in A]; sy = Cint[a.dims[2] for a in A]; sz = Cint[a.dims[3] for a in A]
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
    hist = zeros(Cdouble, cap); it = Ref{Cint}(0); nc = Ref{Cint}(0)
    check(c, ccall((:ns3d_pt_solve, LIB), Cint, (Ptr{Cvoid}, P, P, P, Ref{PtParams}, Ref{Cint}, Ptr{Cdouble}, Cint, Ref{Cint}),
                   c.h, Pr.p, dPrdτ.p, ∇V.p, p, it, hist, cap, nc))
    return Int(it[]), hist[1:nc[]]
end

# ---- level 2: the once-per-step groups and the whole step (M:449-477 / G:121-142) ------------------
struct Fields            # ns3d_fields: the 18 arrays in the script's allocation order (M:343-360)
    Pr::P; dPrdtau::P; C::P; C_o::P; txx::P; tyy::P; tzz::P; txy::P; txz::P; tyz::P
    Vx::P; Vy::P; Vz::P; Vx_o::P; Vy_o::P; Vz_o::P; divV::P; Rp::P
end
Fields(a::DevArray...) = Fields((x.p for x in a)...)
struct StepParams        # ns3d_step_params, field for field
    pt::PtParams
    mu::Cdouble; vin::Cdouble
    a2::Cdouble; b2::Cdouble; ox::Cdouble; oy::Cdouble; sinb::Cdouble; cosb::Cdouble
    xco_g::Cdouble; yco_g::Cdouble; lx::Cdouble; ly::Cdouble
    inlet_guard::Cint; reserved::Cint
end
for (jl, sym) in ((:predictor!, :ns3d_predictor), (:corrector!, :ns3d_corrector), (:advect_swap!, :ns3d_advect_swap))
    @eval $jl(c::Ctx, f::Fields, sp::StepParams) =
        check(c, ccall(($(QuoteNode(sym)), LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}), c.h, f, sp))
end
"One time step = predictor!, pt_solve!, corrector!, advect_swap!; returns (iterations, err history)."
function step!(c::Ctx, f::Fields, sp::StepParams)
    cap = sp.pt.niter ÷ max(sp.pt.nchk, 1) + 2
    hist = zeros(Cdouble, cap); it = Ref{Cint}(0); nc = Ref{Cint}(0)
    check(c, ccall((:ns3d_step, LIB), Cint, (Ptr{Cvoid}, Ref{Fields}, Ref{StepParams}, Ref{Cint}, Ptr{Cdouble}, Cint, Ref{Cint}),
                   c.h, f, sp, it, hist, cap, nc))
    return Int(it[]), hist[1:nc[]]
end

end # module
