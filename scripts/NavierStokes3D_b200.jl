# NavierStokes3D_b200.jl -- scripts/NavierStokes3D_multi_gpu.jl of the reference re-pointed at
# libns3d.so.  Same keyword surface and return value as `run_navierstokes3D` (M:287); the
# physics / numerics block is the reference's, the ParallelStencil, ImplicitGlobalGrid and
# MPI.Allreduce call sites are replaced one for one by NS3DNative (julia/NS3DNative.jl).
#
#   mpirun -np N julia -O3 scripts/NavierStokes3D_b200.jl        (one rank per GPU, z-slabs)
#
# NOT EXECUTED BY JULIA in the build container (no Julia there); its text is executed there by the interpreter of
# oracle/jl_shim.py through julia/NS3DNative.jl into libns3d.so (tests/test_julia_shim_exec.py: bit-equal to the
# reference script's own text, fused and level-1 loop, returned interiors).  The Python twin of this file is
# navierstokes3d_b200/driver.py (`run_navierstokes3D`), which makes exactly these calls through
# the same C ABI and is what the parity tests drive.
include(joinpath(@__DIR__, "..", "julia", "NS3DNative.jl"))
using .NS3DNative
using Printf
import MPI

const USE_FUSED_PT = true     # false: the reference's loop, call site by call site (level 1)

@views function run_navierstokes3D(; do_vis=false, do_save=false, do_print=false, nx=255, nt=10)
    MPI.Initialized() || MPI.Init()
    comm = MPI.COMM_WORLD
    # physics (M:290-319) -----------------------------------------------------------------------
    lx, ρ, vin, μ = 1.0, 1000.0, 1.0, 0.001
    psc = ρ * vin^2
    ly, lz = 0.6 * lx, 0.6 * lx
    ox, oy = -0.4 * lx, 0.0 * lx
    g = 1 / Inf^2 * vin^2 / lx
    a2, b2 = (0.05 * lx)^2, (0.05 * lx)^2
    sinβ, cosβ = sincos(0 * π / 6)
    # numerics (M:322-341); init_global_grid(nx,ny,nz) -> z-slabs, one rank per GPU ----------------
    ny, nz = ceil(Int, nx * 0.6), ceil(Int, nx * 0.6)
    ctx = Ctx(parse(Int, get(ENV, "LOCAL_RANK", string(MPI.Comm_rank(comm)))); mode=FAST)
    me, dims = comm_init_mpi!(ctx, MPI, comm)
    coordz = me
    nx_g, ny_g, nz_g = nx, ny, dims[3] * (nz - 2) + 2
    z_g(iz, dz, sz) = (coordz * (nz - 2) + iz - 1) * dz + 0.5 * (nz - sz) * dz     # IGG z_g
    x_g(ix, dx, sx) = (ix - 1) * dx + 0.5 * (nx - sx) * dx                        # dims[1] == 1
    εit = 1e-3
    niter = 50 * max(nx_g, ny_g, nz_g)
    nchk = ny_g - 1
    dx, dy, dz = lx / nx_g, ly / ny_g, lz / nz_g
    dt = min(1 / 4.1 * max(dx, dy, dz)^2 * ρ / μ, 1.0 * max(dx, dy, dz) / vin)
    damp = 2 / nx
    dτ = 1.0 / sqrt(3.1) * max(dx, dy, dz)
    # allocation (M:343-360): the library's allocator instead of @zeros ----------------------------
    Z(a, b, c) = zeros3(ctx, a, b, c)
    Pr, dPrdτ, C, C_o = Z(nx, ny, nz), Z(nx - 2, ny - 2, nz - 2), Z(nx, ny, nz), Z(nx, ny, nz)
    τxx, τyy, τzz = Z(nx, ny, nz), Z(nx, ny, nz), Z(nx, ny, nz)
    τxy, τxz, τyz = Z(nx - 1, ny - 1, nz - 1), Z(nx - 1, ny - 1, nz - 1), Z(nx - 1, ny - 1, nz - 1)
    Vx, Vy, Vz = Z(nx + 1, ny, nz), Z(nx, ny + 1, nz), Z(nx, ny, nz + 1)
    Vx_o, Vy_o, Vz_o = Z(nx + 1, ny, nz), Z(nx, ny + 1, nz), Z(nx, ny, nz + 1)
    ∇V, Rp = Z(nx, ny, nz), Z(nx - 2, ny - 2, nz - 2)
    xco_g = x_g(1, dx, nx) - (lx - dx) / 2
    yco_g = 0.0 - (ly - dy) / 2
    zco_g = z_g(1, dz, nz) - (lz - dz) / 2
    xvo_g = x_g(1, dx, nx + 1) - (lx - dx) / 2
    xve_g = x_g(nx + 1, dx, nx + 1) - (lx - dx) / 2
    # initialisation (M:369-373) ---------------------------------------------------------------------
    Vy_h = zeros(nx, ny + 1, nz); Vy_h[1, :, :] .= vin                # (sic) M:369
    set!(ctx, Vy, Vy_h)
    set!(ctx, Pr, [-(z_g(iz, dz, nz) - dz / 2) * ρ * g for ix = 1:nx, iy = 1:ny, iz = 1:nz])
    update_halo!(ctx, nz, Pr)
    set_cylinder_M!(ctx, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, zco_g, lx, ly, lz, dx, dy, dz)
    update_halo!(ctx, nz, C, Vx, Vy, Vz)
    pt = PtParams(nx, ny, nz, 0, ρ, dt, dτ, damp, dx, dy, dz, εit, ly^2, psc, niter, nchk,
                  xve_g == lx / 2, 0.0, g, 0, 0)
    # action (M:446-477) -------------------------------------------------------------------------------
    for it = 1:nt
        update_τ!(ctx, τxx, τyy, τzz, τxy, τxz, τyz, Vx, Vy, Vz, μ, dx, dy, dz)
        predict_V!(ctx, Vx, Vy, Vz, τxx, τyy, τzz, τxy, τxz, τyz, ρ, g, dt, dx, dy, dz)
        set_cylinder_M!(ctx, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, zco_g, lx, ly, lz, dx, dy, dz)
        update_halo!(ctx, nz, C, Vx, Vy, Vz)
        update_∇V!(ctx, ∇V, Vx, Vy, Vz, dx, dy, dz)
        update_halo!(ctx, nz, ∇V)
        me == 0 && do_print && println("#it = $it")
        if USE_FUSED_PT
            iters, err_evo = pt_solve!(ctx, Pr, dPrdτ, ∇V, pt)          # M:458-471 in fused kernels
            me == 0 && do_print && @info "  #iter = $iters, err = $(isempty(err_evo) ? NaN : err_evo[end])"
        else
            for iter = 1:niter
                update_dPrdτ!(ctx, Pr, dPrdτ, ∇V, ρ, dt, dτ, damp, dx, dy, dz)
                update_Pr!(ctx, Pr, dPrdτ, dτ)
                set_bc_Pr_M!(ctx, Pr, xve_g, lx, 0.0)                   # includes update_halo!(Pr)
                if iter % nchk == 0
                    compute_res!(ctx, Rp, Pr, ∇V, ρ, dt, dx, dy, dz)
                    err = max_g_abs(ctx, Rp) * ly^2 / psc               # max_g: device max + NCCL allreduce
                    (err < εit || !isfinite(err)) && break
                end
            end
        end
        correct_V!(ctx, Vx, Vy, Vz, Pr, dt, ρ, dx, dy, dz)
        set_cylinder_M!(ctx, C, Vx, Vy, Vz, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, zco_g, lx, ly, lz, dx, dy, dz)
        set_bc_Vel_M!(ctx, Vx, Vy, Vz, xvo_g, lx, vin, Pr)              # includes update_halo!(Vx,Vy,Vz)
        copy!(ctx, Vx_o, Vx); copy!(ctx, Vy_o, Vy); copy!(ctx, Vz_o, Vz); copy!(ctx, C_o, C)
        advect!(ctx, Vx, Vx_o, Vy, Vy_o, Vz, Vz_o, C, C_o, dt, dx, dy, dz)
        update_halo!(ctx, nz, Vx, Vy, Vz)
        if do_save && it % 10 == 0 && me == 0                               # M:515-523 (nsave = 10), single rank
            !ispath("./out_save") && mkdir("./out_save")
            for (name, A) in (("C", C), ("Pr", Pr), ("Vx", Vx), ("Vy", Vy), ("Vz", Vz))
                open(@sprintf("out_save/out_%s_v_%04d.bin", name, it ÷ 10), "w") do io
                    write(io, inner32(ctx, A))                              # Float32 conversion on the device
                end
            end
        end
    end
    # return value (M:528-535): gather!(A_inn, A_v) for every field -- the interior planes are packed on
    # each device and travel to rank 0 over NCCL (ns3d_gather_box); Vz, staggered along the split
    # dimension, takes its extra plane from the last rank only
    np_c = fill(nz - 2, dims[3])
    np_z = [nz - 2 + (r == dims[3] - 1 ? 1 : 0) for r in 0:dims[3]-1]
    return gather_inner(ctx, C, me, np_c), gather_inner(ctx, Pr, me, np_c), gather_inner(ctx, Vx, me, np_c),
           gather_inner(ctx, Vy, me, np_c), gather_inner(ctx, Vz, me, np_z)
end

if abspath(PROGRAM_FILE) == @__FILE__
    run_navierstokes3D(do_vis=false, do_save=false, do_print=true, nx=255, nt=10)
end
