# GEMPICB200.jl -- Julia binding of libgempic_b200.so (include/gempic_b200.h).
#
# Drop-in for the particle-mesh hot path of GEMPIC.jl: the types keep the reference's names,
# constructor signatures and field names, the functions keep their names and argument order;
# each method body is one `ccall`.  Host arrays are only borrowed for the duration of a call
# (`ccall` roots `Ref`/`Array` arguments).  Julia is not available in the build image, so this file
# is desk-checked against the header; the Python mirror gempic.jl_b200/api.py drives the same
# entry points under test.
#
#   using GEMPICB200   # instead of `using GEMPIC` for the types below
module GEMPICB200

export OneDGrid, TwoDGrid, ParticleGroup, ParticleMeshCoupling1D, ParticleMeshCoupling2D, Maxwell1DFEM, TwoDMaxwell,
       HamiltonianSplitting2D3V, operatorHp3, compute_rho_from_e!, compute_rhs_from_function, l2projection, charge_density,
       HamiltonianSplitting, HamiltonianSplittingBoris, strang_splitting!, staggering!,
       operatorHp1, operatorHp2, operatorHE, operatorHB, solve_poisson!,
       add_charge!, evaluate, evaluate_multiple, add_current_update_v!, add_charge_pp!, evaluate_pp, add_current_update_v_pp!, b_to_pp,
       compute_e_from_rho!, compute_e_from_j!,
       compute_e_from_b!, compute_b_from_e!, inner_product, l2norm_squared, l2projection!,
       compute_rhs_from_function!, TimeHistoryDiagnostics, write_step!, upload!, download!, save, load!,
       get_x, get_v, get_charge, get_mass, set_x!, set_v!, set_weights!,
       ParticleSampler, CosSumGaussian, SumCosGaussian, LandauDamping, sample!, sample_synthetic!,
       init, init_devices, device_count, comm_unique_id, comm_init, comm_finalize, comm_size, comm_suspend, synchronize,
       push_v_epart!, push_v_bpart!, push_x_accumulate_j!, upload_fields!, sync_fields!, moments, j_dofs

import FileIO            # save / load! write the reference's JLD2 particle dump (GEMPIC.jl depends on FileIO + JLD2)
using Printf: @sprintf

const LIB = get(ENV, "GEMPIC_B200_LIB", joinpath(@__DIR__, "..", "gempic.jl_b200", "libgempic_b200.so"))
const Handle = UInt64

# status -> exception, mirroring the reference: ArgumentError for its `throw(ArgumentError(..))`
# sites (GEMPIC_EINVAL), AssertionError for its `@assert`s (GEMPIC_EASSERT).
function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:gempic_last_error, LIB), Cstring, ()))
    rc == 1 && throw(ArgumentError(msg))
    rc == 2 && throw(AssertionError(msg))
    error("gempic_b200 status $rc: $msg")
end

const _initialised = Ref(false)
function init(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
    _initialised[] && return
    check(ccall((:gempic_init, LIB), Cint, (Cint,), device))
    _initialised[] = true
    atexit(() -> ccall((:gempic_finalize, LIB), Cint, ()))
end
# ONE Julia process driving several GPUs (the reference chunks its particles over Julia threads inside one process,
# src/hamiltonian_splitting.jl:61-66): call this instead of `init`, before any object is created.  Everything below then
# acts on all devices -- ParticleGroup{D,V}(n) takes the global particle count, `pg.array` is the global array,
# strang_splitting! on 1e9 particles over 8 GPUs is a single `ccall`.
function init_devices(devices::AbstractVector{<:Integer})
    _initialised[] && throw(ArgumentError("GEMPICB200 is already initialised"))
    ids = convert(Vector{Cint}, devices)
    check(ccall((:gempic_init_devices, LIB), Cint, (Cint, Ptr{Cint}), length(ids), ids))
    _initialised[] = true
    atexit(() -> ccall((:gempic_finalize, LIB), Cint, ()))
end
init_devices(n::Integer) = init_devices(collect(0:(n - 1)))
device_count() = Int(ccall((:gempic_device_count, LIB), Cint, ()))
# One process per GPU instead (MPI.jl, Distributed, ...): rank 0 makes the id, the host side broadcasts its 128 bytes,
# every rank calls comm_init after init(local_device).  Replaces reduce(+, fetch.(tasks)) (hamiltonian_splitting_1d2v.jl:88).
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:gempic_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init(n_ranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:gempic_comm_init, LIB), Cint, (Cint, Cint, Ptr{UInt8}), n_ranks, rank, id))
comm_finalize() = check(ccall((:gempic_comm_finalize, LIB), Cint, ()))
comm_size() = Int(ccall((:gempic_comm_size, LIB), Cint, ()))
comm_suspend(on::Bool) = check(ccall((:gempic_comm_suspend, LIB), Cint, (Cint,), on ? 1 : 0))

# ---- mesh (src/mesh.jl:52-67): only the scalars are read by the hot path --------------------------
struct OneDGrid
    xmin::Float64
    xmax::Float64
    nx::Int
    dimx::Float64
    OneDGrid(xmin, xmax, nx) = new(xmin, xmax, nx, xmax - xmin)
end

# ---- ParticleGroup{D,V} (src/particle_group.jl:15-46) -------------------------------------------
# The particles live on the device.  `pg.array` is a lazily synchronised host mirror in the reference layout
# (D+V+W) x N (SURVEY section 8b), handed out as a write-tracking wrapper (MirrorArray): element READS download the rows
# first if the device has advanced since the last download and leave the mirror clean; element WRITES do the same and
# mark it as newer, so the next device operation uploads it.  Because the wrapper looks at the group's state on every
# access, an alias kept across device calls (`a = pg.array; strang_splitting!(...); a[2, 1]`) stays coherent, and a
# loop that only reads `pg.array` never re-uploads anything.  A 1e9-particle run that never touches `pg.array` never
# copies anything (host_mirror = false does not even allocate it); `upload!` / `download!` remain for explicit control.
mutable struct ParticleGroup{D,V}
    dims::Tuple{Int,Int}
    n_particles::Int
    host::Array{Float64,2}
    host_newer::Bool
    dev_newer::Bool
    common_weight::Float64
    charge::Float64
    mass::Float64
    n_weights::Int
    q_over_m::Float64
    handle::Handle
    function ParticleGroup{D,V}(n_particles; charge = 1.0, mass = 1.0, n_weights = 1, common_weight = 0.0,
                                host_mirror = true) where {D,V}
        init()
        common_weight == 0.0 && (common_weight = 1.0 / n_particles)
        h = Ref{Handle}(0)
        check(ccall((:gempic_pg_create, LIB), Cint,
                    (Cint, Cint, Cint, Int64, Cdouble, Cdouble, Cdouble, Ref{Handle}),
                    D, V, n_weights, n_particles, charge, mass, common_weight, h))
        host = host_mirror ? zeros(Float64, D + V + n_weights, n_particles) : zeros(Float64, D + V + n_weights, 0)
        pg = new(( D, V ), n_particles, host, false, false, common_weight, charge, mass, n_weights, charge / mass, h[])
        finalizer(p -> ccall((:gempic_pg_destroy, LIB), Cint, (Handle,), p.handle), pg)
        return pg
    end
end
function upload!(pg::ParticleGroup)
    host = getfield(pg, :host)
    size(host, 2) == pg.n_particles || throw(ArgumentError("ParticleGroup: the host mirror holds $(size(host, 2)) of $(pg.n_particles) particles"))
    check(ccall((:gempic_pg_upload, LIB), Cint, (Handle, Ptr{Cdouble}), pg.handle, host))
    setfield!(pg, :host_newer, false)
    setfield!(pg, :dev_newer, false)
    return nothing
end
function download!(pg::ParticleGroup)
    size(getfield(pg, :host), 2) == pg.n_particles ||
        setfield!(pg, :host, zeros(Float64, sum(pg.dims) + pg.n_weights, pg.n_particles))
    host = getfield(pg, :host)
    check(ccall((:gempic_pg_download, LIB), Cint, (Handle, Ptr{Cdouble}), pg.handle, host))
    setfield!(pg, :dev_newer, false)
    return host
end
struct MirrorArray{D,V} <: AbstractMatrix{Float64}
    pg::ParticleGroup{D,V}
end
Base.size(a::MirrorArray) = (sum(getfield(a.pg, :dims)) + getfield(a.pg, :n_weights), getfield(a.pg, :n_particles))
Base.IndexStyle(::Type{<:MirrorArray}) = IndexCartesian()
Base.@propagate_inbounds Base.getindex(a::MirrorArray, i::Int, j::Int) = _mirror(a.pg)[i, j]
Base.@propagate_inbounds function Base.setindex!(a::MirrorArray, v, i::Int, j::Int)
    host = _mirror(a.pg)
    setfield!(a.pg, :host_newer, true)
    host[i, j] = v
    return a
end
Base.parent(a::MirrorArray) = _mirror(a.pg)      # the plain Array (reads only: writes through it are not tracked)
function Base.getproperty(pg::ParticleGroup, name::Symbol)
    name === :array || return getfield(pg, name)
    return MirrorArray(pg)
end
function Base.setproperty!(pg::ParticleGroup, name::Symbol, value)
    name === :array || return setfield!(pg, name, convert(fieldtype(typeof(pg), name), value))
    setfield!(pg, :host, convert(Array{Float64,2}, value))
    setfield!(pg, :host_newer, true)
    setfield!(pg, :dev_newer, false)
    return value
end
# read-only view of the mirror (the reference's get_x / get_v / get_charge / get_mass): synchronised, not marked as written
function _mirror(pg::ParticleGroup)
    (getfield(pg, :dev_newer) || size(getfield(pg, :host), 2) != getfield(pg, :n_particles)) && download!(pg)
    return getfield(pg, :host)
end
# before a device operation on the particles: push a host mirror the caller may have written; after one that changes them
_flush(pg::ParticleGroup) = (getfield(pg, :host_newer) && upload!(pg); nothing)
_touched(pg::ParticleGroup) = (setfield!(pg, :dev_newer, true); nothing)
# save (src/particle_group.jl:152-165): the reference's JLD2 dump, from a fresh download of the device rows
function save(file, step, p::ParticleGroup{D,V}) where {D,V}
    _flush(p)
    a = _mirror(p)
    datafile = @sprintf("%s-%06d.jld2", file, step)
    FileIO.save(datafile, Dict("x" => a[1:D, :], "v" => a[(D + 1):(D + V), :], "w" => a[(D + V + 1):end, :]))
end
# restart: fill the device rows from such a dump
function load!(p::ParticleGroup{D,V}, datafile) where {D,V}
    d = FileIO.load(datafile)
    p.array = vcat(d["x"], d["v"], d["w"])
    _flush(p)
end
get_x(p::ParticleGroup{D,V}, i::Int) where {D,V} = _mirror(p)[1:D, i]                    # src/particle_group.jl:53
get_v(p::ParticleGroup{D,V}, i::Int) where {D,V} = _mirror(p)[(D + 1):(D + V), i]        # :60
get_charge(p::ParticleGroup{D,V}, i::Int; i_wi = 1) where {D,V} = p.charge * _mirror(p)[D + V + i_wi, i] * p.common_weight   # :67-69
get_mass(p::ParticleGroup{D,V}, i::Int; i_wi = 1) where {D,V} = p.mass * _mirror(p)[D + V + i_wi, i] * p.common_weight       # :76-78
# set_x! / set_v! / set_weights! (:94-150): write the mirror; the next device operation uploads it
set_x!(p::ParticleGroup{D,V}, i::Int, x::Vector{Float64}) where {D,V} = (p.array[1:D, i] .= x[1:D]; nothing)
set_x!(p::ParticleGroup{D,V}, i::Int, x::Float64) where {D,V} = (p.array[1, i] = x; nothing)
set_v!(p::ParticleGroup{D,V}, i::Int, v::Vector{Float64}) where {D,V} = (p.array[(D + 1):(D + V), i] .= v[1:V]; nothing)
set_v!(p::ParticleGroup{D,V}, i::Int, v::Float64) where {D,V} = (p.array[D + 1, i] = v; nothing)
set_weights!(p::ParticleGroup{D,V}, i::Int, w::Vector{Float64}) where {D,V} = (p.array[(D + V + 1):(D + V + p.n_weights), i] .= w; nothing)
set_weights!(p::ParticleGroup{D,V}, i::Int, w::Float64) where {D,V} = (p.array[D + V + 1, i] = w; nothing)
function sort!(pg::ParticleGroup, pmc)
    _flush(pg)
    check(ccall((:gempic_pg_sort, LIB), Cint, (Handle, Handle), pg.handle, pmc.handle))
    _touched(pg)
end

# ---- ParticleMeshCoupling1D (src/particle_mesh_coupling_1d.jl:26-95) ----------------------------
mutable struct ParticleMeshCoupling1D
    dims::Int
    xmin::Float64
    Lx::Float64
    delta_x::Float64
    n_grid::Int
    n_dofs::Int
    no_particles::Int
    spline_degree::Int
    n_span::Int
    scaling::Float64
    handle::Handle
    function ParticleMeshCoupling1D(mesh::OneDGrid, no_particles, spline_degree, smoothing_type::Symbol)
        init()
        smoothing_type in (:collocation, :galerkin) ||
            throw(ArgumentError("Smoothing Type $smoothing_type not implemented for kernel_smoother_spline_1d. "))
        h = Ref{Handle}(0)
        check(ccall((:gempic_pmc1d_create, LIB), Cint, (Cdouble, Cdouble, Cint, Int64, Cint, Cint, Ref{Handle}),
                    mesh.xmin, mesh.xmax, mesh.nx, no_particles, spline_degree, smoothing_type == :galerkin ? 1 : 0, h))
        dx = (mesh.xmax - mesh.xmin) / mesh.nx
        p = new(1, mesh.xmin, mesh.xmax - mesh.xmin, dx, mesh.nx, mesh.nx, no_particles, spline_degree,
                spline_degree + 1, smoothing_type == :collocation ? 1 / dx : 1.0, h[])
        finalizer(q -> ccall((:gempic_pmc1d_destroy, LIB), Cint, (Handle,), q.handle), p)
        return p
    end
end

# batched forms of the per-particle methods (:261-280, :438-453, :296-376, :471-529)
function add_charge!(rho_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, position::Vector{Float64}, marker_charge::Vector{Float64})
    check(ccall((:gempic_pmc1d_add_charge, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                p.handle, position, marker_charge, length(position), rho_dofs))
end
add_charge!(rho_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, position::Float64, marker_charge::Float64) =
    add_charge!(rho_dofs, p, [position], [marker_charge])
function add_charge!(rho_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, pg::ParticleGroup)
    _flush(pg)
    check(ccall((:gempic_pmc1d_add_charge_pg, LIB), Cint, (Handle, Handle, Ptr{Cdouble}), p.handle, pg.handle, rho_dofs))
end
function evaluate(p::ParticleMeshCoupling1D, position::Vector{Float64}, field_dofs::Vector{Float64})
    out = similar(position)
    check(ccall((:gempic_pmc1d_evaluate, LIB), Cint, (Handle, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                p.handle, position, length(position), field_dofs, out))
    return out
end
evaluate(p::ParticleMeshCoupling1D, position::Float64, field_dofs::Vector{Float64}) = evaluate(p, [position], field_dofs)[1]
function add_current_update_v!(j_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, position_old::Vector{Float64},
                               position_new::Vector{Float64}, marker_charge::Vector{Float64}, qoverm::Float64,
                               bfield_dofs::Vector{Float64}, vi::Vector{Float64})
    check(ccall((:gempic_pmc1d_add_current_update_v, LIB), Cint,
                (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                p.handle, position_old, position_new, marker_charge, qoverm, bfield_dofs, vi, length(vi), j_dofs))
    return vi
end

function add_current_update_v!(j_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, position_old::Float64, position_new::Float64,
                               marker_charge::Float64, qoverm::Float64, bfield_dofs::Vector{Float64}, vi::Float64)
    return add_current_update_v!(j_dofs, p, [position_old], [position_new], [marker_charge], qoverm, bfield_dofs, [vi])[1]
end
# 1d1v form without B (:471-529)
function add_current_update_v!(j_dofs::Vector{Float64}, p::ParticleMeshCoupling1D, position_old::Vector{Float64},
                               position_new::Vector{Float64}, marker_charge::Vector{Float64})
    check(ccall((:gempic_pmc1d_add_current, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                p.handle, position_old, position_new, marker_charge, length(position_old), j_dofs))
end
function evaluate(p::ParticleMeshCoupling1D, pg::ParticleGroup, field_dofs::Vector{Float64})
    _flush(pg)
    out = zeros(Float64, pg.n_particles)
    check(ccall((:gempic_pmc1d_evaluate_pg, LIB), Cint, (Handle, Handle, Ptr{Cdouble}, Ptr{Cdouble}), p.handle, pg.handle, field_dofs, out))
    return out
end
# The `_pp` forms of the reference (:106-250) compute the same functions through the piecewise-polynomial tables; the
# kernels always evaluate in pp form, so they are aliases.  b_to_pp (src/splinepp.jl:241-261) returns an opaque handle on
# the dofs: the device builds the cell polynomials itself.
struct PPField
    dofs::Vector{Float64}
end
b_to_pp(p::ParticleMeshCoupling1D, field_dofs::Vector{Float64}) = PPField(copy(field_dofs))
add_charge_pp!(rho_dofs, p::ParticleMeshCoupling1D, position, marker_charge) = add_charge!(rho_dofs, p, position, marker_charge)
evaluate_pp(p::ParticleMeshCoupling1D, position, field_dofs_pp::PPField) = evaluate(p, position, field_dofs_pp.dofs)
add_current_update_v_pp!(j_dofs, p::ParticleMeshCoupling1D, args...) = add_current_update_v!(j_dofs, p, args...)

# ---- Maxwell1DFEM (src/maxwell_1d_fem.jl:29-177) ------------------------------------------------
mutable struct Maxwell1DFEM
    Lx::Float64
    xmin::Float64
    delta_x::Float64
    n_dofs::Int
    s_deg_0::Int
    s_deg_1::Int
    handle::Handle
    function Maxwell1DFEM(mesh::OneDGrid, degree::Int)
        init()
        h = Ref{Handle}(0)
        check(ccall((:gempic_maxwell1d_create, LIB), Cint, (Cdouble, Cdouble, Cint, Cint, Ref{Handle}),
                    mesh.xmin, mesh.xmax, mesh.nx, degree, h))
        m = new(mesh.xmax - mesh.xmin, mesh.xmin, (mesh.xmax - mesh.xmin) / mesh.nx, mesh.nx, degree, degree - 1, h[])
        finalizer(q -> ccall((:gempic_maxwell1d_destroy, LIB), Cint, (Handle,), q.handle), m)
        return m
    end
end
compute_e_from_rho!(e::Vector{Float64}, m::Maxwell1DFEM, rho::Vector{Float64}) =
    check(ccall((:gempic_maxwell1d_compute_e_from_rho, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}), m.handle, e, rho))
compute_e_from_j!(e::Vector{Float64}, m::Maxwell1DFEM, j::Vector{Float64}, component::Int) =
    check(ccall((:gempic_maxwell1d_compute_e_from_j, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Cint), m.handle, e, j, component))
compute_e_from_b!(e::Vector{Float64}, m::Maxwell1DFEM, dt::Float64, b::Vector{Float64}) =
    check(ccall((:gempic_maxwell1d_compute_e_from_b, LIB), Cint, (Handle, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}), m.handle, e, dt, b))
compute_b_from_e!(b::Vector{Float64}, m::Maxwell1DFEM, dt::Float64, e::Vector{Float64}) =
    check(ccall((:gempic_maxwell1d_compute_b_from_e, LIB), Cint, (Handle, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}), m.handle, b, dt, e))
function inner_product(m::Maxwell1DFEM, c1::Vector{Float64}, c2::Vector{Float64}, degree)
    out = Ref{Cdouble}(0)
    check(ccall((:gempic_maxwell1d_inner_product, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ref{Cdouble}),
                m.handle, c1, c2, degree, out))
    return out[]
end
l2norm_squared(m::Maxwell1DFEM, c::Vector{Float64}, degree) = inner_product(m, c, c, degree)

# C callback trampoline for the set-up helpers that take a Julia function of x
_tramp(x::Cdouble, ctx::Ptr{Cvoid})::Cdouble = unsafe_pointer_to_objref(ctx)[](x)
function l2projection!(coefs::Vector{Float64}, m::Maxwell1DFEM, f::Function, degree)
    r = Ref{Function}(f)
    GC.@preserve r check(ccall((:gempic_maxwell1d_l2projection, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                               m.handle, coefs, @cfunction(_tramp, Cdouble, (Cdouble, Ptr{Cvoid})), pointer_from_objref(r), degree))
end
function compute_rhs_from_function!(coefs::Vector{Float64}, m::Maxwell1DFEM, f::Function, degree)
    r = Ref{Function}(f)
    GC.@preserve r check(ccall((:gempic_maxwell1d_compute_rhs_from_function, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                               m.handle, coefs, @cfunction(_tramp, Cdouble, (Cdouble, Ptr{Cvoid})), pointer_from_objref(r), degree))
end

# ---- HamiltonianSplitting{D,V} (src/hamiltonian_splitting.jl:20-86) -----------------------------
# e_dofs / b_dofs stay the caller's arrays (aliased, :80-81): every call copies them in and the
# updated values back (3 x n doubles), which is what the reference's tests read after each operator.
# j_dofs (:51) is scratch owned by the splitting object; no caller of the reference reads it.  It stays on the device and
# `h.j_dofs` reads it back on access (getproperty below), so a strang_splitting! call moves 3 x n doubles each way.
mutable struct HamiltonianSplitting{D,V}
    dims::Tuple{Int64,Int64}
    maxwell_solver::Maxwell1DFEM
    kernel_smoother_0::ParticleMeshCoupling1D
    kernel_smoother_1::ParticleMeshCoupling1D
    particle_group::ParticleGroup
    e_dofs::Array{Array{Float64,1}}
    b_dofs::Array{Float64,1}
    j_host::Array{Array{Float64,1}}
    handle::Handle
    resident::Bool   # true: the fields stay on the device between calls (upload_fields! / sync_fields! move them explicitly)
    function HamiltonianSplitting{D,V}(maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group,
                                       e_dofs, b_dofs; fuse = true, resident = false) where {D,V}
        h = Ref{Handle}(0)
        check(ccall((:gempic_hs_create, LIB), Cint, (Cint, Cint, Handle, Handle, Handle, Handle, Ref{Handle}),
                    D, V, maxwell_solver.handle, kernel_smoother_0.handle, kernel_smoother_1.handle, particle_group.handle, h))
        check(ccall((:gempic_hs_set_fusion, LIB), Cint, (Handle, Cint), h[], fuse ? 1 : 0))
        j_dofs = [zeros(Float64, kernel_smoother_0.n_dofs) for i in 1:2]
        hs = new((D, V), maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group, e_dofs, b_dofs, j_dofs, h[], resident)
        # the library keeps what a splitting points to alive until it is destroyed, so finalizers may run in any order
        finalizer(q -> ccall((:gempic_hs_destroy, LIB), Cint, (Handle,), getfield(q, :handle)), hs)
        resident && upload_fields!(hs)
        return hs
    end
end
function Base.getproperty(h::HamiltonianSplitting, name::Symbol)
    name === :j_dofs || return getfield(h, name)
    j = getfield(h, :j_host)   # after a fused strang_splitting! this rebuilds j_dofs[2] from the particles (DESIGN.md section 6)
    check(ccall((:gempic_hs_get_fields, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                getfield(h, :handle), C_NULL, C_NULL, C_NULL, j[1], j[2]))
    return j
end
upload_fields!(h::HamiltonianSplitting) =
    check(ccall((:gempic_hs_set_fields, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.handle, h.e_dofs[1], h.e_dofs[2], h.b_dofs))
sync_fields!(h::HamiltonianSplitting) =
    check(ccall((:gempic_hs_get_fields, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, h.e_dofs[1], h.e_dofs[2], h.b_dofs, C_NULL, C_NULL))
synchronize() = check(ccall((:gempic_synchronize, LIB), Cint, ()))
const OP_HP1, OP_HP2, OP_HE, OP_HB = Cint(1), Cint(2), Cint(3), Cint(4)
function _op(h::HamiltonianSplitting, op::Cint, dt::Float64)
    _flush(h.particle_group)
    if h.resident
        check(ccall((:gempic_hs_operator, LIB), Cint, (Handle, Cint, Cdouble), h.handle, op, dt))
    else
        check(ccall((:gempic_hs_operator_host, LIB), Cint,
                    (Handle, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    h.handle, op, dt, h.e_dofs[1], h.e_dofs[2], h.b_dofs, C_NULL, C_NULL))
    end
    _touched(h.particle_group)
end
operatorHp1(h::HamiltonianSplitting, dt::Float64) = _op(h, OP_HP1, dt)   # src/hamiltonian_splitting_1d2v.jl:41-112 / _1d1v.jl:63-98
operatorHp2(h::HamiltonianSplitting, dt::Float64) = _op(h, OP_HP2, dt)   # :129-176
operatorHE(h::HamiltonianSplitting, dt::Float64) = _op(h, OP_HE, dt)     # :191-219
operatorHB(h::HamiltonianSplitting, dt::Float64) = _op(h, OP_HB, dt)     # :234-236 / _1d1v.jl:113-125
function strang_splitting!(h::HamiltonianSplitting, dt::Float64, number_steps::Int)   # src/hamiltonian_splitting.jl:98-108
    _flush(h.particle_group)
    if h.resident
        check(ccall((:gempic_hs_strang_splitting, LIB), Cint, (Handle, Cdouble, Int64), h.handle, dt, number_steps))
    else
        check(ccall((:gempic_hs_strang_splitting_host, LIB), Cint,
                    (Handle, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    h.handle, dt, number_steps, h.e_dofs[1], h.e_dofs[2], h.b_dofs, C_NULL, C_NULL))
    end
    _touched(h.particle_group)
end

# ---- HamiltonianSplittingBoris (src/hamiltonian_splitting_boris.jl:23-88) -----------------------
mutable struct HamiltonianSplittingBoris
    maxwell_solver::Maxwell1DFEM
    kernel_smoother_0::ParticleMeshCoupling1D
    kernel_smoother_1::ParticleMeshCoupling1D
    particle_group::ParticleGroup
    e_dofs::Array{Array{Float64,1}}
    b_dofs::Array{Float64,1}
    handle::Handle
    function HamiltonianSplittingBoris(maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group, e_dofs, b_dofs)
        h = Ref{Handle}(0)
        check(ccall((:gempic_boris_create, LIB), Cint, (Handle, Handle, Handle, Handle, Ref{Handle}),
                    maxwell_solver.handle, kernel_smoother_0.handle, kernel_smoother_1.handle, particle_group.handle, h))
        hs = new(maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group, e_dofs, b_dofs, h[])
        finalizer(q -> ccall((:gempic_boris_destroy, LIB), Cint, (Handle,), q.handle), hs)
        return hs
    end
end
# e_dofs_mid / b_dofs_mid / j_dofs of the reference struct (:30-36) stay on the device; read one back with boris_field
function boris_field(h::HamiltonianSplittingBoris, which::Integer)   # GEMPIC_F_E1 .. GEMPIC_F_B_MID = 0 .. 7
    out = zeros(Float64, h.kernel_smoother_0.n_dofs)
    _flush(h.particle_group)
    check(ccall((:gempic_boris_get_field, LIB), Cint, (Handle, Cint, Ptr{Cdouble}), h.handle, which, out))
    return out
end
function staggering!(h::HamiltonianSplittingBoris, dt::Float64)                     # :99-122
    _flush(h.particle_group)
    check(ccall((:gempic_boris_staggering_host, LIB), Cint, (Handle, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, dt, h.e_dofs[1], h.e_dofs[2], h.b_dofs))
    _touched(h.particle_group)
end
# the separate pushes of the reference (:189-288) on the staggered fields the object holds; device-resident variants
# of staggering! / strang_splitting! for callers that keep the fields on the device (boris_upload_fields! first)
function _push(h::HamiltonianSplittingBoris, sym::Symbol, dt::Float64)
    _flush(h.particle_group)
    rc = sym === :e ? ccall((:gempic_boris_push_v_epart, LIB), Cint, (Handle, Cdouble), h.handle, dt) :
         sym === :b ? ccall((:gempic_boris_push_v_bpart, LIB), Cint, (Handle, Cdouble), h.handle, dt) :
                      ccall((:gempic_boris_push_x_accumulate_j, LIB), Cint, (Handle, Cdouble), h.handle, dt)
    check(rc)
    _touched(h.particle_group)
end
push_v_epart!(h::HamiltonianSplittingBoris, dt::Float64) = _push(h, :e, dt)            # :189-204
push_v_bpart!(h::HamiltonianSplittingBoris, dt::Float64) = _push(h, :b, dt)            # :211-233
push_x_accumulate_j!(h::HamiltonianSplittingBoris, dt::Float64) = _push(h, :x, dt)     # :250-288
boris_upload_fields!(h::HamiltonianSplittingBoris) =
    check(ccall((:gempic_boris_set_fields, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.handle, h.e_dofs[1], h.e_dofs[2], h.b_dofs))
function staggering_resident!(h::HamiltonianSplittingBoris, dt::Float64)
    _flush(h.particle_group)
    check(ccall((:gempic_boris_staggering, LIB), Cint, (Handle, Cdouble), h.handle, dt))
    _touched(h.particle_group)
end
function strang_splitting_resident!(h::HamiltonianSplittingBoris, dt::Float64, number_steps::Int)
    _flush(h.particle_group)
    check(ccall((:gempic_boris_strang_splitting, LIB), Cint, (Handle, Cdouble, Int64), h.handle, dt, number_steps))
    _touched(h.particle_group)
end
function strang_splitting!(h::HamiltonianSplittingBoris, dt::Float64, number_steps::Int)   # :132-177
    _flush(h.particle_group)
    check(ccall((:gempic_boris_strang_splitting_host, LIB), Cint, (Handle, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, dt, number_steps, h.e_dofs[1], h.e_dofs[2], h.b_dofs))
    _touched(h.particle_group)
end

# ---- TwoDGrid / TwoDMaxwell (src/mesh.jl:17-45, src/maxwell_2d_fem.jl:11-87) --------------------
struct TwoDGrid
    nx::Int
    ny::Int
    xmin::Float64
    xmax::Float64
    ymin::Float64
    ymax::Float64
    TwoDGrid(xmin, xmax, nx::Int, ymin, ymax, ny::Int) = new(nx, ny, xmin, xmax, ymin, ymax)
end
mutable struct TwoDMaxwell
    s_deg_0::Int
    s_deg_1::Int
    mesh::TwoDGrid
    handle::Handle
    function TwoDMaxwell(mesh::TwoDGrid, degree::Int)
        init()
        h = Ref{Handle}(0)
        check(ccall((:gempic_maxwell2d_create, LIB), Cint, (Cdouble, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Ref{Handle}),
                    mesh.xmin, mesh.xmax, mesh.nx, mesh.ymin, mesh.ymax, mesh.ny, degree, h))
        m = new(degree, degree - 1, mesh, h[])
        finalizer(q -> ccall((:gempic_maxwell2d_destroy, LIB), Cint, (Handle,), q.handle), m)
        return m
    end
end
const Dofs3 = Vector{Vector{Float64}}
compute_e_from_rho!(e::Dofs3, m::TwoDMaxwell, rho::Vector{Float64}) =                      # :199-201
    check(ccall((:gempic_maxwell2d_compute_e_from_rho, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), m.handle, e[1], e[2], rho))
compute_e_from_b!(e::Dofs3, m::TwoDMaxwell, dt::Float64, b::Dofs3) =                       # :370-412
    check(ccall((:gempic_maxwell2d_compute_e_from_b, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                m.handle, e[1], e[2], e[3], dt, b[1], b[2], b[3]))
compute_b_from_e!(b::Dofs3, m::TwoDMaxwell, dt::Float64, e::Dofs3) =                       # :423-444
    check(ccall((:gempic_maxwell2d_compute_b_from_e, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                m.handle, b[1], b[2], b[3], dt, e[1], e[2], e[3]))
compute_e_from_j!(e::Vector{Float64}, m::TwoDMaxwell, current::Vector{Float64}, component::Int) =   # :455-459
    check(ccall((:gempic_maxwell2d_compute_e_from_j, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Cint), m.handle, e, current, component))
compute_rho_from_e!(rho::Vector{Float64}, m::TwoDMaxwell, e::Dofs3) =                       # :468-500
    check(ccall((:gempic_maxwell2d_compute_rho_from_e, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), m.handle, rho, e[1], e[2], e[3]))
function inner_product(m::TwoDMaxwell, c1::Vector{Float64}, c2::Vector{Float64}, component, form)   # :512-575
    out = Ref{Cdouble}(0)
    check(ccall((:gempic_maxwell2d_inner_product, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ref{Cdouble}),
                m.handle, c1, c2, component, form, out))
    return out[]
end
_tramp2(x::Cdouble, y::Cdouble, ctx::Ptr{Cvoid})::Cdouble = unsafe_pointer_to_objref(ctx)[](x, y)
function compute_rhs_from_function(m::TwoDMaxwell, f::Function, component::Int, form::Int)   # :124-196
    coefs = zeros(m.mesh.nx * m.mesh.ny)
    r = Ref{Function}(f)
    GC.@preserve r check(ccall((:gempic_maxwell2d_compute_rhs_from_function, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                               m.handle, coefs, @cfunction(_tramp2, Cdouble, (Cdouble, Cdouble, Ptr{Cvoid})), pointer_from_objref(r), component, form))
    return coefs
end
function l2projection(m::TwoDMaxwell, f::Function, component::Int, form::Int)                # :283-293
    coefs = zeros(m.mesh.nx * m.mesh.ny)
    r = Ref{Function}(f)
    GC.@preserve r check(ccall((:gempic_maxwell2d_l2projection, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                               m.handle, coefs, @cfunction(_tramp2, Cdouble, (Cdouble, Cdouble, Ptr{Cvoid})), pointer_from_objref(r), component, form))
    return coefs
end

# ---- HamiltonianSplitting{2,3}: the content of the reference's empty src/hamiltonian_splitting_2d3v.jl ----
# Same field names and operator methods as HamiltonianSplitting{1,2}; e_dofs / b_dofs are three aliased
# nx*ny vectors each.  strang_splitting! = HB HE Hp3 Hp2 Hp1 Hp2 Hp3 HE HB.
mutable struct HamiltonianSplitting2D3V
    dims::Tuple{Int64,Int64}
    maxwell_solver::TwoDMaxwell
    particle_group::ParticleGroup
    e_dofs::Dofs3
    b_dofs::Dofs3
    handle::Handle
    function HamiltonianSplitting2D3V(maxwell_solver::TwoDMaxwell, particle_group::ParticleGroup{2,3}, e_dofs::Dofs3, b_dofs::Dofs3)
        h = Ref{Handle}(0)
        check(ccall((:gempic_hs2d_create, LIB), Cint, (Handle, Handle, Ref{Handle}), maxwell_solver.handle, particle_group.handle, h))
        hs = new((2, 3), maxwell_solver, particle_group, e_dofs, b_dofs, h[])
        finalizer(q -> ccall((:gempic_hs2d_destroy, LIB), Cint, (Handle,), q.handle), hs)
        return hs
    end
end
const OP_HP3 = Cint(5)
function _op(h::HamiltonianSplitting2D3V, op::Cint, dt::Float64)
    _flush(h.particle_group)
    check(ccall((:gempic_hs2d_operator_host, LIB), Cint, (Handle, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, op, dt, h.e_dofs[1], h.e_dofs[2], h.e_dofs[3], h.b_dofs[1], h.b_dofs[2], h.b_dofs[3]))
    _touched(h.particle_group)
end
operatorHp1(h::HamiltonianSplitting2D3V, dt::Float64) = _op(h, OP_HP1, dt)
operatorHp2(h::HamiltonianSplitting2D3V, dt::Float64) = _op(h, OP_HP2, dt)
operatorHp3(h::HamiltonianSplitting2D3V, dt::Float64) = _op(h, OP_HP3, dt)
operatorHE(h::HamiltonianSplitting2D3V, dt::Float64) = _op(h, OP_HE, dt)
operatorHB(h::HamiltonianSplitting2D3V, dt::Float64) = _op(h, OP_HB, dt)
function strang_splitting!(h::HamiltonianSplitting2D3V, dt::Float64, number_steps::Int)
    _flush(h.particle_group)
    check(ccall((:gempic_hs2d_strang_splitting_host, LIB), Cint, (Handle, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, dt, number_steps, h.e_dofs[1], h.e_dofs[2], h.e_dofs[3], h.b_dofs[1], h.b_dofs[2], h.b_dofs[3]))
    _touched(h.particle_group)
end
# device-resident forms (fields stay on the GPU between calls) and the current dofs of the last Hp1, Hp2, Hp3
upload_fields!(h::HamiltonianSplitting2D3V) =
    check(ccall((:gempic_hs2d_set_fields, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, h.e_dofs[1], h.e_dofs[2], h.e_dofs[3], h.b_dofs[1], h.b_dofs[2], h.b_dofs[3]))
sync_fields!(h::HamiltonianSplitting2D3V) =
    check(ccall((:gempic_hs2d_get_fields, LIB), Cint,
                (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, h.e_dofs[1], h.e_dofs[2], h.e_dofs[3], h.b_dofs[1], h.b_dofs[2], h.b_dofs[3], C_NULL, C_NULL, C_NULL))
function j_dofs(h::HamiltonianSplitting2D3V)
    n = h.maxwell_solver.mesh.nx * h.maxwell_solver.mesh.ny
    j = [zeros(Float64, n) for c in 1:3]
    check(ccall((:gempic_hs2d_get_fields, LIB), Cint,
                (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                h.handle, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, j[1], j[2], j[3]))
    return j
end
function operator_resident!(h::HamiltonianSplitting2D3V, op::Cint, dt::Float64)
    _flush(h.particle_group)
    check(ccall((:gempic_hs2d_operator, LIB), Cint, (Handle, Cint, Cdouble), h.handle, op, dt))
    _touched(h.particle_group)
end
function strang_splitting_resident!(h::HamiltonianSplitting2D3V, dt::Float64, number_steps::Int)
    _flush(h.particle_group)
    check(ccall((:gempic_hs2d_strang_splitting, LIB), Cint, (Handle, Cdouble, Int64), h.handle, dt, number_steps))
    _touched(h.particle_group)
end
function charge_density(h::HamiltonianSplitting2D3V)
    _flush(h.particle_group)
    rho = zeros(h.maxwell_solver.mesh.nx * h.maxwell_solver.mesh.ny)
    check(ccall((:gempic_hs2d_charge_density, LIB), Cint, (Handle, Ptr{Cdouble}), h.handle, rho))
    return rho
end

function moments(h::HamiltonianSplitting2D3V)   # sum_p w |v|^2, sum_p w v_k (diagnostics.jl:197-211 in three velocity dimensions)
    _flush(h.particle_group)
    out = zeros(Float64, 4)
    check(ccall((:gempic_hs2d_moments, LIB), Cint, (Handle, Ptr{Cdouble}), h.handle, out))
    return out
end
set_sort_interval!(h::HamiltonianSplitting2D3V, interval::Integer) =
    check(ccall((:gempic_hs2d_set_sort_interval, LIB), Cint, (Handle, Cint), h.handle, interval))
set_fusion!(h::HamiltonianSplitting2D3V, level::Integer) = check(ccall((:gempic_hs2d_set_fusion, LIB), Cint, (Handle, Cint), h.handle, level))
function sort!(pg::ParticleGroup{2,V}, m::TwoDMaxwell) where {V}
    _flush(pg)
    check(ccall((:gempic_pg_sort2d, LIB), Cint, (Handle, Handle), pg.handle, m.handle))
    _touched(pg)
end

# ---- ParticleMeshCoupling2D (src/particle_mesh_coupling_2d.jl:12-231) ------------------------------------------
mutable struct ParticleMeshCoupling2D
    grid::TwoDGrid
    npart::Int
    spline_degree::Int
    n_span::Int
    scaling::Float64
    n_dofs::Int
    handle::Handle
    function ParticleMeshCoupling2D(pg::ParticleGroup{D,V}, grid::TwoDGrid, degree::Int, smoothing_type::Symbol) where {D,V}
        init()
        smoothing_type in (:collocation, :galerkin) ||
            throw(ArgumentError("Smoothing Type $smoothing_type not implemented for kernel_smoother_spline_2d. "))
        h = Ref{Handle}(0)
        check(ccall((:gempic_pmc2d_create, LIB), Cint, (Cdouble, Cdouble, Cint, Cdouble, Cdouble, Cint, Cint, Cint, Ref{Handle}),
                    grid.xmin, grid.xmax, grid.nx, grid.ymin, grid.ymax, grid.ny, degree, smoothing_type == :galerkin ? 1 : 0, h))
        dx, dy = (grid.xmax - grid.xmin) / grid.nx, (grid.ymax - grid.ymin) / grid.ny
        p = new(grid, pg.n_particles, degree, degree + 1, smoothing_type == :collocation ? 1 / (dx * dy) : 1.0, grid.nx * grid.ny, h[])
        finalizer(q -> ccall((:gempic_pmc2d_destroy, LIB), Cint, (Handle,), q.handle), p)
        return p
    end
end
# add_charge!(rho_dofs, pm, xp, yp, wp) (:189-205): scalars or equal-length vectors
function add_charge!(rho_dofs::Vector{Float64}, pm::ParticleMeshCoupling2D, xp::Vector{Float64}, yp::Vector{Float64}, wp::Vector{Float64})
    check(ccall((:gempic_pmc2d_add_charge, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                pm.handle, xp, yp, wp, length(xp), rho_dofs))
end
add_charge!(rho_dofs::Vector{Float64}, pm::ParticleMeshCoupling2D, xp::Float64, yp::Float64, wp::Float64) =
    add_charge!(rho_dofs, pm, [xp], [yp], [wp])
function add_charge!(rho_dofs::Vector{Float64}, pm::ParticleMeshCoupling2D, pg::ParticleGroup)
    _flush(pg)
    check(ccall((:gempic_pmc2d_add_charge_pg, LIB), Cint, (Handle, Handle, Ptr{Cdouble}), pm.handle, pg.handle, rho_dofs))
end
# evaluate(pm, xp, yp, field_dofs) (:207-221)
function evaluate(pm::ParticleMeshCoupling2D, xp::Vector{Float64}, yp::Vector{Float64}, field_dofs::Vector{Float64})
    out = similar(xp)
    check(ccall((:gempic_pmc2d_evaluate, LIB), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                pm.handle, xp, yp, length(xp), field_dofs, out))
    return out
end
evaluate(pm::ParticleMeshCoupling2D, xp::Float64, yp::Float64, field_dofs::Vector{Float64}) = evaluate(pm, [xp], [yp], field_dofs)[1]
function evaluate(pm::ParticleMeshCoupling2D, pg::ParticleGroup, field_dofs::Vector{Float64})
    _flush(pg)
    out = zeros(Float64, pg.n_particles)
    check(ccall((:gempic_pmc2d_evaluate_pg, LIB), Cint, (Handle, Handle, Ptr{Cdouble}, Ptr{Cdouble}), pm.handle, pg.handle, field_dofs, out))
    return out
end
# evaluate_multiple(pm, position, field_dofs) (:223-231): two fields at once
function evaluate_multiple(pm::ParticleMeshCoupling2D, xp::Vector{Float64}, yp::Vector{Float64}, field_dofs)
    o1, o2 = similar(xp), similar(xp)
    check(ccall((:gempic_pmc2d_evaluate_multiple, LIB), Cint,
                (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pm.handle, xp, yp, length(xp), field_dofs[1], field_dofs[2], o1, o2))
    return o1, o2
end
function evaluate_multiple(pm::ParticleMeshCoupling2D, position, field_dofs)
    o1, o2 = evaluate_multiple(pm, [Float64(position[1])], [Float64(position[2])], field_dofs)
    return o1[1], o2[1]
end
# `_pp` entry points (:107-188): same functions through the pp tables (floor instead of ceil cell index; identical
# values by continuity of the splines)
b_to_pp(pm::ParticleMeshCoupling2D, field_dofs::Vector{Float64}) = PPField(copy(field_dofs))
add_charge_pp!(rho_dofs, pm::ParticleMeshCoupling2D, xp, yp, wp) = add_charge!(rho_dofs, pm, xp, yp, wp)
evaluate_pp(pm::ParticleMeshCoupling2D, xp, yp, pp::PPField) = evaluate(pm, xp, yp, pp.dofs)

# ---- samplers (src/particle_sampling.jl, src/distributions.jl, src/landau_damping.jl) on the device --------------
struct CosGaussianParams
    dims::Tuple{Int64,Int64}
    n_cos::Int64
    n_gaussians::Int64
    k::Array{Vector{Float64},1}
    α::Vector{Float64}
    σ::Array{Vector{Float64},1}
    μ::Array{Vector{Float64},1}
    normal::Vector{Float64}
    δ::Vector{Float64}
    function CosGaussianParams(dims, k, α, σ, μ, δ = [1.0])          # distributions.jl:24-47
        n_cos = length(k)
        @assert n_cos == length(α)
        for i in 1:n_cos
            @assert length(k[i]) == dims[1]
        end
        n_gaussians = length(σ)
        @assert n_gaussians == length(μ)
        @assert all([all(s .!= 0.0) for s in σ])
        normal = [1.0 / ((2π)^(0.5 * dims[2]) * prod(σ[j])) for j in 1:n_gaussians]
        @assert sum(δ) == 1.0
        return new(dims, n_cos, n_gaussians, k, α, σ, μ, normal, δ)
    end
end
abstract type AbstractCosGaussian end
struct CosSumGaussian{D,V} <: AbstractCosGaussian                    # distributions.jl:88-107
    dims::Tuple{Int64,Int64}
    params::CosGaussianParams
    CosSumGaussian{D,V}(k, α, σ, μ, δ = [1.0]) where {D,V} = new((D, V), CosGaussianParams((D, V), k, α, σ, μ, δ))
end
struct SumCosGaussian{D,V} <: AbstractCosGaussian                    # distributions.jl:142-157
    dims::Tuple{Int64,Int64}
    params::CosGaussianParams
    SumCosGaussian{D,V}(k, α, σ, μ, δ = [1.0]) where {D,V} = new((D, V), CosGaussianParams((D, V), k, α, σ, μ, δ))
end
struct LandauDamping                                                 # landau_damping.jl:10-13
    alpha::Float64
    kx::Float64
end
struct ParticleSampler{D,V}                                          # particle_sampling.jl:14-59
    sampling_type::Symbol
    dims::Tuple{Int,Int}
    n_particles::Int
    symmetric::Bool
    seed::Int
    function ParticleSampler{D,V}(sampling_type::Symbol, symmetric::Bool, n_particles::Int, seed::Int = 1234) where {D,V}
        sampling_type in (:random, :sobol) || throw(ArgumentError("Sampling type $sampling_type not implemented"))
        if symmetric
            np = mod(n_particles, 2^(D + V))
            np != 0 && (n_particles += np)                           # (sic) :35-38
        end
        return new(sampling_type, (D, V), n_particles, symmetric, seed)
    end
    ParticleSampler{D,V}(sampling_type::Symbol, n_particles::Int, seed::Int = 1234) where {D,V} =
        ParticleSampler{D,V}(sampling_type, false, n_particles, seed)
end
function _sample_landau!(pg::ParticleGroup, α, k, σ, weight; first_index = 0, n_global = pg.n_particles)
    check(ccall((:gempic_pg_sample_landau, LIB), Cint, (Handle, Cdouble, Cdouble, Cdouble, Cdouble, Int64, Int64),
                pg.handle, α, k, σ, weight, first_index, n_global))
    setfield!(pg, :host_newer, false)
    _touched(pg)
end
# sample!(pg, α, k, σ, mesh) (particle_sampling.jl:266-311): Sobol(2) + Newton, deterministic; `first_index` / `n_global`
# for a group that holds a shard of a larger load (one process per GPU)
sample!(pg::ParticleGroup{1,1}, α::Float64, k::Float64, σ::Float64, mesh::OneDGrid; kw...) = _sample_landau!(pg, α, k, σ, mesh.dimx; kw...)
function sample!(pg::ParticleGroup{1,2}, α::Float64, k::Float64, σ::Float64, mesh::OneDGrid; kw...)
    @assert mesh.dimx ≈ 2π / k
    _sample_landau!(pg, α, k, σ, mesh.dimx; kw...)
end
sample!(pg::ParticleGroup{1,1}, ps::ParticleSampler, df::AbstractCosGaussian, mesh::OneDGrid; kw...) =      # :248-256
    sample!(pg, df.params.α[1], df.params.k[1][1], df.params.σ[1][1], mesh; kw...)
sample!(d::LandauDamping, pg::ParticleGroup{1,2}; first_index = 0, n_global = pg.n_particles) =               # landau_damping.jl:34-59
    _sample_landau!(pg, d.alpha, d.kx, 1.0, 2π / d.kx / n_global; first_index = first_index, n_global = n_global)
# sample!(pg, ps, df, mesh) (:68-225): sample_all / sample_sym; Sobol coordinates as Sobol.jl's, normal deviates from the
# library's counter-based generator (statistical parity with the reference's MersenneTwister draws)
function sample!(pg::ParticleGroup{1,2}, ps::ParticleSampler, df::AbstractCosGaussian, mesh::OneDGrid; first_index = 0)
    p = df.params
    k = Float64[kk[1] for kk in p.k]
    σ = Float64[p.σ[j][c] for j in 1:p.n_gaussians for c in 1:2]
    μ = Float64[p.μ[j][c] for j in 1:p.n_gaussians for c in 1:2]
    check(ccall((:gempic_pg_sample_cos_gaussian, LIB), Cint,
                (Handle, Cint, Cint, UInt64, Cdouble, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64),
                pg.handle, ps.sampling_type == :sobol ? 1 : 0, ps.symmetric ? 1 : 0, UInt64(ps.seed), mesh.xmin, mesh.dimx,
                p.n_cos, k, p.α, p.n_gaussians, σ, μ, p.δ, first_index))
    setfield!(pg, :host_newer, false)
    _touched(pg)
end
# synthetic benchmark loads of the library (kind :uniform or :landau, counter-based generator)
function sample_synthetic!(pg::ParticleGroup, kind::Symbol, xmin, L; alpha = 0.0, k = 1.0, sigma = [1.0, 1.0, 1.0], seed = 1234, first_index = 0)
    sig = convert(Vector{Float64}, vcat(sigma, [1.0, 1.0, 1.0])[1:3])
    check(ccall((:gempic_pg_sample, LIB), Cint, (Handle, Cint, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, UInt64, Int64),
                pg.handle, kind == :landau ? 1 : 0, xmin, L, alpha, k, sig, UInt64(seed), first_index))
    setfield!(pg, :host_newer, false)
    _touched(pg)
end

# ---- diagnostics (src/diagnostics.jl) --------------------------------------------------------------
function solve_poisson!(efield::Vector{Float64}, particle_group::ParticleGroup, kernel_smoother_0::ParticleMeshCoupling1D,
                        maxwell_solver::Maxwell1DFEM, rho::Vector{Float64})            # :15-31
    _flush(particle_group)
    check(ccall((:gempic_solve_poisson, LIB), Cint, (Handle, Handle, Handle, Ptr{Cdouble}, Ptr{Cdouble}),
                particle_group.handle, kernel_smoother_0.handle, maxwell_solver.handle, efield, rho))
end
# TimeHistoryDiagnostics (:127-170).  `data` holds one NamedTuple per write_step! with the reference's eleven columns
# (:143-155) -- a Tables.jl row table, so `DataFrame(thdiag.data)`, `thdiag.data[end].KineticEnergy` and
# `getproperty.(thdiag.data, :PotentialEnergyE1)` all work without making DataFrames a dependency of the shim.
const DiagRow = NamedTuple{(:Time, :KineticEnergy, :Momentum1, :Momentum2, :PotentialEnergyE1, :PotentialEnergyE2,
                            :PotentialEnergyB3, :Transfer, :VVB, :Poynting, :ErrorPoisson),NTuple{11,Float64}}
struct TimeHistoryDiagnostics
    particle_group::ParticleGroup
    maxwell_solver::Maxwell1DFEM
    kernel_smoother_0::ParticleMeshCoupling1D
    kernel_smoother_1::ParticleMeshCoupling1D
    data::Vector{DiagRow}
    TimeHistoryDiagnostics(particle_group::ParticleGroup, maxwell_solver::Maxwell1DFEM, kernel_smoother_0::ParticleMeshCoupling1D,
                           kernel_smoother_1::ParticleMeshCoupling1D) =
        new(particle_group, maxwell_solver, kernel_smoother_0, kernel_smoother_1, DiagRow[])
end
# write_step!(thdiag, time, degree, efield_dofs, bfield_dofs, efield_dofs_n, efield_poisson) (:186-250): the particle sums and
# the field products are taken on the device (one pass; none at all right after a fused strang_splitting! + solve_poisson!,
# DESIGN.md section 6) and the row is pushed onto thdiag.data like the reference does
function write_step!(thdiag::TimeHistoryDiagnostics, time, degree, efield_dofs, bfield_dofs, efield_dofs_n, efield_poisson)
    out = zeros(Float64, 11)
    pg = thdiag.particle_group
    _flush(pg)
    check(ccall((:gempic_diag_write_step, LIB), Cint,
                (Handle, Handle, Handle, Handle, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pg.handle, thdiag.maxwell_solver.handle, thdiag.kernel_smoother_0.handle, thdiag.kernel_smoother_1.handle, time, degree,
                efield_dofs[1], efield_dofs[2], bfield_dofs, efield_dofs_n[1], efield_dofs_n[2], efield_poisson, out))
    return push!(thdiag.data, DiagRow(Tuple(out)))
end

end # module
