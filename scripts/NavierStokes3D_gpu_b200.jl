# NavierStokes3D_gpu_b200.jl -- scripts/NavierStokes3D_gpu.jl (G) of the reference re-pointed at libns3d.so.
#
# The text of `runme` is the reference's (G:12-173); what changes is the head of the file: instead of
# ParallelStencil (`@init_parallel_stencil`, `@zeros`, `@parallel`, `Data.Array`) the look-alike surface of
# julia/NS3DNative.jl is loaded, whose macros act on a default context of the library.  Three spellings differ,
# because a device array of the library is not a broadcastable CuArray:
#     Vx_o .= Vx                      ->  copy!(Vx_o, Vx)                       (G:141)
#     maximum(abs.(Rp))               ->  max_g(abs, Rp)                        (G:132; one pass, NaN-propagating)
#     Data.Array([... comprehension]) ->  Data.Array(collect(...)) of a 3-D Array{Float64}   (G:85-88)
# `USE_FUSED = true` replaces the body of the pseudo-transient loop (G:126-137) by the fused entry point
# `pt_solve!`, which returns the same iteration count and residual history.
#
#   julia -O3 scripts/NavierStokes3D_gpu_b200.jl
#
# NOT EXECUTED BY JULIA in the build container (no Julia there); its text is executed there by the interpreter of
# oracle/jl_shim.py through julia/NS3DNative.jl into libns3d.so (tests/test_julia_shim_exec.py: bit-equal to the
# reference script's own text with USE_FUSED = true and false).  The Python twin of this file is
# navierstokes3d_b200/driver.py (`runme`), which makes exactly these calls through the same C ABI and is what the
# parity tests drive (tests/test_gpu_solver.py::test_config_B_whole_time_steps_vs_oracle runs this configuration
# for two time steps, bit-exact against the CPU oracle).
include(joinpath(@__DIR__, "..", "julia", "NS3DNative.jl"))
using .NS3DNative
using LinearAlgebra, Printf
@init_ns3d(0, FAST)                     # stands where @init_parallel_stencil(CUDA, Float64, 3) stood (G:5)

const USE_FUSED = true
const set_cylinder! = set_cylinder_G!   # G:336-368

# set_bc_Vel! / set_bc_Pr! as in the script (G:264-286): the face kernels one by one
function set_bc_Vel!(Vx, Vy, Vz, Vprof)
    @parallel (1:size(Vx,2), 1:size(Vx,3)) bc_x!(Vx)
    @parallel (1:size(Vx,1), 1:size(Vx,3)) bc_y!(Vx)
    @parallel (1:size(Vx,1), 1:size(Vx,2)) bc_zV!(Vx)
    @parallel (1:size(Vy,2), 1:size(Vy,3)) bc_x!(Vy)
    @parallel (1:size(Vy,1), 1:size(Vy,3)) bc_y!(Vy)
    @parallel (1:size(Vy,1), 1:size(Vy,2)) bc_zV!(Vy)
    @parallel (1:size(Vz,2), 1:size(Vz,3)) bc_x!(Vz)
    @parallel (1:size(Vz,1), 1:size(Vz,3)) bc_y!(Vz)
    @parallel (1:size(Vz,1), 1:size(Vz,2)) bc_zV!(Vz)
    return
end
function set_bc_Pr!(Pr, dz, nz, g, ρ)
    @parallel (1:size(Pr,1), 1:size(Pr,3)) bc_y!(Pr)
    @parallel (1:size(Pr,1), 1:size(Pr,2)) bc_z!(Pr)
    @parallel (1:size(Pr,2), 1:size(Pr,3)) bc_xhydstatic!(Pr, dz, nz, g, ρ)
    return
end

@views function runme(; do_vis=false, do_save=false, nt=10000)
    # physics (G:15-41)
    lx        = 1.0
    ρ         = 1000.0
    vin       = 1.0
    μ         = 0.001
    psc       = ρ*vin^2
    ly_lx     = 0.6
    lz_lx     = 0.6
    a_lx      = 0.05
    b_lx      = 0.05
    ox_lx     = -0.3
    oy_lx     = 0.0
    β         = 0*π/6
    ly        = ly_lx*lx
    lz        = lz_lx*lx
    ox        = ox_lx*lx
    oy        = oy_lx*lx
    g         = 9.81
    a2        = (a_lx*lx)^2
    b2        = (b_lx*lx)^2
    sinβ,cosβ = sincos(β)
    # numerics (G:44-56)
    nx        = 255
    ny        = ceil(Int,nx*ly_lx)
    nz        = ceil(Int,nx*lz_lx)
    εit       = 1e-3
    niter     = 50*max(ny,nz)
    nchk      = 1*(ny-1)
    nsave     = 10
    CFLτ      = 1.0/sqrt(3.1)
    CFL_visc  = 1/4.1
    CFL_adv   = 1.0
    # preprocessing (G:58-63)
    dx,dy,dz  = lx/nx,ly/ny,lz/nz
    dt        = min(CFL_visc*max(dx,dy,dz)^2*ρ/μ,CFL_adv*max(dx,dy,dz)/vin)
    damp      = 2/nx
    dτ        = CFLτ*max(dx,dy,dz)
    xc,yc,zc  = LinRange(-(lx-dx)/2,(lx-dx)/2,nx  ),LinRange(-(ly-dy)/2,(ly-dy)/2,ny  ),LinRange(-(lz-dz)/2,(lz-dz)/2,nz  )
    xv,yv,zv  = LinRange(-lx/2     ,lx/2     ,nx+1),LinRange(-ly/2     ,ly/2     ,ny+1),LinRange(-lz/2     ,lz/2     ,nz+1)
    # allocation (G:65-82)
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
    # init (G:85-89)
    Vprof     = [vin*(7.0/6.0)*((zc[iz]+lz/2)/lz)^(1.0/6.0) for iz=1:nz]
    Vx        = Data.Array(collect(Float64, [vin*(7.0/6.0)*((zc[iz]+lz/2)/lz)^(1.0/6.0) + 0*yc[iy] + 0*xv[ix]  for ix=1:(nx+1),iy=1:ny,iz=1:nz]))
    Pr        = Data.Array(collect(Float64, [-(zc[iz]-lz/2)*ρ*g + 0*yc[iy] + 0*xc[ix] for ix=1:nx,iy=1:ny,iz=1:nz]))
    pt        = PtParams(nx, ny, nz, VARIANT_G, ρ, dt, dτ, damp, dx, dy, dz, εit, ly^2, psc, niter, nchk, 0, 0.0, g, 0, 0)
    # action (G:119-142)
    for it = 1:nt
        err_evo = Float64[]; iter_evo = Float64[]
        @parallel update_τ!(τxx,τyy,τzz,τxy,τxz,τyz,Vx,Vy,Vz,μ,dx,dy,dz)
        @parallel predict_V!(Vx,Vy,Vz,τxx,τyy,τzz,τxy,τxz,τyz,ρ,g,dt,dx,dy,dz)
        @parallel set_cylinder!(C,Vx,Vy,Vz,a2,b2,ox,oy,sinβ,cosβ,lx,ly,lz,dx,dy,dz)
        @parallel update_∇V!(∇V,Vx,Vy,Vz,dx,dy,dz)
        println("#it = $it")
        if USE_FUSED
            iters, err_evo = @parallel pt_solve!(Pr,dPrdτ,∇V,pt)      # G:126-137 in fused kernels
            for (c, err) in enumerate(err_evo)
                @printf("  #iter = %d, err = %1.3e\n", min(c*nchk, iters), err)
            end
        else
            for iter = 1:niter
                @parallel update_dPrdτ!(Pr,dPrdτ,∇V,ρ,dt,dτ,damp,dx,dy,dz)
                @parallel update_Pr!(Pr,dPrdτ,dτ)
                set_bc_Pr!(Pr, dz, nz, g, ρ)
                if iter % nchk == 0
                    @parallel compute_res!(Rp,Pr,∇V,ρ,dt,dx,dy,dz)
                    err = max_g(abs, Rp)*ly^2/psc
                    push!(err_evo, err); push!(iter_evo,iter/ny)
                    @printf("  #iter = %d, err = %1.3e\n", iter, err)
                    if err < εit || !isfinite(err) break end
                end
            end
        end
        @parallel correct_V!(Vx,Vy,Vz,Pr,dt,ρ,dx,dy,dz)
        @parallel set_cylinder!(C,Vx,Vy,Vz,a2,b2,ox,oy,sinβ,cosβ,lx,ly,lz,dx,dy,dz)
        set_bc_Vel!(Vx, Vy, Vz, Vprof)
        copy!(Vx_o, Vx); copy!(Vy_o, Vy); copy!(Vz_o, Vz); copy!(C_o, C)
        @parallel advect!(Vx,Vx_o,Vy,Vy_o,Vz,Vz_o,C,C_o,dt,dx,dy,dz)
        if do_save && it % nsave == 0                                  # G:168-170 (needs MAT.jl)
            !ispath("./out_save") && mkdir("./out_save")
            Main.MAT.matwrite("out_save/step_$it.mat",Dict("Pr"=>Array(Pr),"Vx"=>Array(Vx),"Vy"=>Array(Vy),"Vz"=>Array(Vz),"C"=>Array(C),"dx"=>dx,"dy"=>dy,"dz"=>dz))
        end
    end
    return
end

if abspath(PROGRAM_FILE) == @__FILE__
    runme(; do_vis=false, do_save=false)
end
