# NavierStokes3D_multi_gpu_b200.jl -- scripts/NavierStokes3D_multi_gpu.jl (M) of the reference with its TEXT KEPT and only
# its head re-pointed at libns3d.so: instead of ParallelStencil / ImplicitGlobalGrid / MPI.Allreduce the look-alike surface
# of julia/NS3DNative.jl is loaded (`@zeros`, `@parallel`, `Data.Array`, `init_global_grid`, `nx_g`, `x_g`, `update_halo!`,
# `max_g`, `gather!`, `finalize_global_grid`), acting on the library's default context.  (scripts/NavierStokes3D_b200.jl is
# the same run written against the explicit-context API.)  Spellings that differ from M, because a device array of the
# library is neither indexable from the host nor broadcastable:
#     Vy[1,:,:] .= vin                ->  fill_plane_x!(default_ctx(), Vy, 1, vin)        (M:369)
#     Data.Array([... comprehension]) ->  Data.Array(collect(Float64, [...]))            (M:370)
#     max_g(abs.(Rp))                 ->  max_g(abs, Rp)                                  (M:466; one pass, NaN-propagating)
#     Vx_o .= Vx                      ->  copy!(Vx_o, Vx)                                 (M:475)
#     C_inn .= Array(C)[2:end-1,...]; gather!(C_inn, C_v)  ->  gather!(C, C_v)            (M:399-403, 528-532: the interior
#                                                                                        is extracted on the device)
#     init_global_grid(nx, ny, nz)    ->  init_global_grid(nx, ny, nz; MPI=MPI)           (z-slabs, one rank per GPU)
# `USE_FUSED = true` replaces the body of the pseudo-transient loop (M:458-471) by `pt_solve!`.  Plotting is left out.
#
#   mpirun -np N julia -O3 scripts/NavierStokes3D_multi_gpu_b200.jl
#
# NOT EXECUTED BY JULIA in the build container (no Julia there); its text is executed there by the interpreter of
# oracle/jl_shim.py through julia/NS3DNative.jl into libns3d.so on one rank (tests/test_julia_shim_exec.py: bit-equal to
# the reference script's own text, USE_FUSED true and false, incl. the returned interiors).
include(joinpath(@__DIR__, "..", "julia", "NS3DNative.jl"))
using .NS3DNative
using Printf
import MPI

const USE_FUSED = true
const set_cylinder! = set_cylinder_M!   # M:249-281

# set_bc_Vel! / set_bc_Pr! as in the script (M:156-184): face kernels, the float == guards, the halo update
function set_bc_Vel!(Vx, Vy, Vz, xvo_g, lx, vin)
    @parallel (1:size(Vx,2),1:size(Vx,3)) bc_x!(Vx)
    @parallel (1:size(Vx,1),1:size(Vx,3)) bc_y!(Vx)
    @parallel (1:size(Vx,1),1:size(Vx,2)) bc_z!(Vx)
    @parallel (1:size(Vy,2),1:size(Vy,3)) bc_x!(Vy)
    @parallel (1:size(Vy,1),1:size(Vy,2)) bc_z!(Vy)
    @parallel (1:size(Vz,2),1:size(Vz,3)) bc_x!(Vz)
    @parallel (1:size(Vz,1),1:size(Vz,3)) bc_y!(Vz)
    if xvo_g == -lx/2
        @parallel (1:size(Vx,2),1:size(Vx,3)) bc_x_Vx!(Vx, vin)
    end
    update_halo!(Vx,Vy,Vz)
    return nothing
end
function set_bc_Pr!(Pr, xve_g, lx, val)
    @parallel (1:size(Pr,2),1:size(Pr,3)) bc_x!(Pr)
    @parallel (1:size(Pr,1),1:size(Pr,3)) bc_y!(Pr)
    @parallel (1:size(Pr,1),1:size(Pr,2)) bc_z!(Pr)
    if xve_g == lx/2
        @parallel (1:size(Pr,2),1:size(Pr,3)) bc_x_Pr!(Pr, val)
    end
    update_halo!(Pr)
    return nothing
end

@views function run_navierstokes3D(; do_vis=false,do_save=false,do_print=false,nx=255,nt=10)
    # physics (M:290-319)
    lx        = 1.0
    ρ         = 1000.0
    vin       = 1.0
    μ         = 0.001
    psc       = ρ*vin^2
    Fr        = Inf
    ly_lx     = 0.6
    lz_lx     = 0.6
    a_lx      = 0.05
    b_lx      = 0.05
    ox_lx     = -0.4
    oy_lx     = 0.0
    β         = 0*π/6
    ly        = ly_lx*lx
    lz        = lz_lx*lx
    ox        = ox_lx*lx
    oy        = oy_lx*lx
    g         = 1/Fr^2*vin^2/lx
    a2        = (a_lx*lx)^2
    b2        = (b_lx*lx)^2
    sinβ,cosβ = sincos(β)
    # numerics (M:322-335)
    ny        = ceil(Int,nx*ly_lx)
    nz        = ceil(Int,nx*lz_lx)
    me, dims  = init_global_grid(nx, ny, nz; MPI=MPI)
    εit       = 1e-3
    niter     = 50*max(nx_g(),ny_g(),nz_g())
    nchk      = 1*(ny_g()-1)
    CFLτ      = 1.0/sqrt(3.1)
    CFL_visc  = 1/4.1
    CFL_adv   = 1.0
    # preprocessing (M:338-341)
    dx,dy,dz  = lx/nx_g(),ly/ny_g(),lz/nz_g()
    dt        = min(CFL_visc*max(dx,dy,dz)^2*ρ/μ,CFL_adv*max(dx,dy,dz)/vin)
    damp      = 2/nx
    dτ        = CFLτ*max(dx,dy,dz)
    # allocation (M:343-361)
    Pr        = @zeros(nx  ,ny  ,nz  )
    dPrdτ     = @zeros(nx-2,ny-2,nz-2)
    C         = @zeros(nx  ,ny  ,nz  )
    C_o       = @zeros(nx  ,ny  ,nz  )
    τxx       = @zeros(nx  ,ny  ,nz  )
    τyy       = @zeros(nx  ,ny  ,nz  )
    τzz       = @zeros(nx  ,ny  ,nz  )
    τxy       = @zeros(nx-1,ny-1,nz-1)
    τxz       = @zeros(nx-1,ny-1,nz-1)
    τyz       = @zeros(nx-1,ny-1,nz-1)
    Vx        = @zeros(nx+1,ny  ,nz  )
    Vy        = @zeros(nx  ,ny+1,nz  )
    Vz        = @zeros(nx  ,ny  ,nz+1)
    Vx_o      = @zeros(nx+1,ny  ,nz  )
    Vy_o      = @zeros(nx  ,ny+1,nz  )
    Vz_o      = @zeros(nx  ,ny  ,nz+1)
    ∇V        = @zeros(nx  ,ny  ,nz  )
    Rp        = @zeros(nx-2,ny-2,nz-2)
    xc,yc,zc  = LinRange(-(lx-dx)/2,(lx-dx)/2,nx),LinRange(-(ly-dy)/2,(ly-dy)/2,ny),LinRange(-(lz-dz)/2,(lz-dz)/2,nz)
    # global coordinates for initial and boundary conditions (M:363-367)
    xco_g     = x_g(1   ,dx,C ) - (lx-dx)/2
    yco_g     = y_g(1   ,dy,C ) - (ly-dy)/2
    zco_g     = z_g(1   ,dz,C ) - (lz-dz)/2
    xvo_g     = x_g(1   ,dx,Vx) - (lx-dx)/2
    xve_g     = x_g(nx+1,dx,Vx)- (lx-dx)/2
    # initialization (M:369-373)
    fill_plane_x!(default_ctx(), Vy, 1, vin)
    Pr         = Data.Array(collect(Float64, [-(z_g(iz,dz,C )-dz/2)*ρ*g + 0*yc[iy] + 0*zc[iz] for ix=1:size(C ,1),iy=1:size(C ,2),iz=1:size(C ,3)]))
    update_halo!(Pr)
    @parallel set_cylinder!(C,Vx,Vy,Vz,a2,b2,ox,oy,sinβ,cosβ,xco_g,yco_g,zco_g,lx,ly,lz,dx,dy,dz)
    update_halo!(C,Vx,Vy,Vz)
    pt        = PtParams(nx, ny, nz, VARIANT_M, ρ, dt, dτ, damp, dx, dy, dz, εit, ly^2, psc, niter, nchk, xve_g == lx/2, 0.0, g, 0, 0)
    # global arrays for the return value (M:378-390)
    nx_v,ny_v,nz_v = (nx-2)*dims[1],(ny-2)*dims[2],(nz-2)*dims[3]
    C_v    = zeros(nx_v  , ny_v  , nz_v  )
    Pr_v   = zeros(nx_v  , ny_v  , nz_v  )
    Vx_v   = zeros(nx_v+1, ny_v  , nz_v  )
    Vy_v   = zeros(nx_v  , ny_v+1, nz_v  )
    Vz_v   = zeros(nx_v  , ny_v  , nz_v+1)
    # action (M:446-477)
    for it = 1:nt
        err_evo = Float64[]; iter_evo = Float64[]
        @parallel update_τ!(τxx,τyy,τzz,τxy,τxz,τyz,Vx,Vy,Vz,μ,dx,dy,dz)
        update_halo!(τxx,τyy,τzz)
        @parallel predict_V!(Vx,Vy,Vz,τxx,τyy,τzz,τxy,τxz,τyz,ρ,g,dt,dx,dy,dz)
        @parallel set_cylinder!(C,Vx,Vy,Vz,a2,b2,ox,oy,sinβ,cosβ,xco_g,yco_g,zco_g,lx,ly,lz,dx,dy,dz)
        update_halo!(C,Vx,Vy,Vz)
        @parallel update_∇V!(∇V,Vx,Vy,Vz,dx,dy,dz)
        update_halo!(∇V)
        if me==0 if do_print  println("#it = $it") end end
        if USE_FUSED
            iters, err_evo = @parallel pt_solve!(Pr,dPrdτ,∇V,pt)      # M:458-471 in fused kernels, halos inside
        else
            for iter = 1:niter
                @parallel update_dPrdτ!(Pr,dPrdτ,∇V,ρ,dt,dτ,damp,dx,dy,dz)
                update_halo!(∇V)
                @parallel update_Pr!(Pr,dPrdτ,dτ)
                update_halo!(Pr)
                set_bc_Pr!(Pr, xve_g, lx, 0.0)
                if iter % nchk == 0
                    @parallel compute_res!(Rp,Pr,∇V,ρ,dt,dx,dy,dz)
                    err = max_g(abs, Rp)*ly^2/psc
                    push!(err_evo, err); push!(iter_evo,iter/ny_g())
                    if me==0 if do_print @printf("  #iter = %d, err = %1.3e\n", iter, err) end end
                    if err < εit || !isfinite(err) break end
                end
            end
        end
        @parallel correct_V!(Vx,Vy,Vz,Pr,dt,ρ,dx,dy,dz)
        @parallel set_cylinder!(C,Vx,Vy,Vz,a2,b2,ox,oy,sinβ,cosβ,xco_g,yco_g,zco_g,lx,ly,lz,dx,dy,dz)
        set_bc_Vel!(Vx, Vy, Vz, xvo_g, lx, vin)
        copy!(Vx_o, Vx); copy!(Vy_o, Vy); copy!(Vz_o, Vz); copy!(C_o, C)
        @parallel advect!(Vx,Vx_o,Vy,Vy_o,Vz,Vz_o,C,C_o,dt,dx,dy,dz)
        update_halo!(Vx,Vy,Vz)
    end
    # gather the interiors for the return call (M:528-535)
    gather!(C , C_v )
    gather!(Pr, Pr_v)
    gather!(Vx, Vx_v)
    gather!(Vy, Vy_v)
    gather!(Vz, Vz_v)
    finalize_global_grid()
    return C_v,Pr_v,Vx_v,Vy_v,Vz_v
end

if abspath(PROGRAM_FILE) == @__FILE__
    run_navierstokes3D(do_vis=false, do_save=false, do_print=true, nx=255, nt=10)
end
