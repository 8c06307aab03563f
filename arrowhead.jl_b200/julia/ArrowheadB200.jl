# ArrowheadB200.jl -- the reference-side binding of libarrowhead_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` (see DESIGN.md section 1).  It is the
# file a maintainer of ivanslapnicar/Arrowhead.jl would `include` after `src/arrowhead6.jl`; it replaces
# exactly the per-k loop bodies of the three drivers and leaves every other line (sorting, deflation, givens,
# the final views) untouched:
#
#   solve!(Λ,U,κ,Ax::SymArrow,zx)          <->  src/arrowhead3.jl:627-632 and :639-644
#   solve!(Λ,U,κ,Ax::SymDPR1,zx)           <->  src/arrowhead4.jl:516-522 and :529-535
#   solve!(Σ,U,V,κ,Ax::HalfArrow,zx,zxv,signD) <-> src/arrowhead6.jl:622-628 and :635-642
#
# The Python mirror of the same calls (`arrowhead.jl_b200/drivers.py`) is what the test-suite exercises.  Because
# Julia cannot run here, the struct layout this file relies on is pinned another way: `tests/test_abi_layout.py`
# compiles a plain-C harness against include/arrowhead_b200.h, prints sizeof/offsetof of every field, and compares
# them with the ctypes mirror AND with the offsets Julia's C-compatible struct layout rules give for `AhResult`
# below (field order and types are parsed out of this very file).
module ArrowheadB200

using LinearAlgebra
import ..Arrowhead: SymArrow, SymDPR1, HalfArrow, Info

const LIB = get(ENV, "ARROWHEAD_B200_LIB", "libarrowhead_b200.so")
const AH_MEM_HOST   = Int32(0)
const AH_MEM_DEVICE = Int32(1)

# status word: low byte = conditions, bits 8..14 = branch trace (include/arrowhead_b200.h)
const AH_ST_NU_INF         = Int32(1)
const AH_ST_NEEDS_BIGFLOAT = Int32(2)    # the reference switches to BigFloat here (Qout 50/100)
const AH_ST_POLE_RETURN    = Int32(4)    # Remedy 3 came back with a pole
const AH_ST_REF_THROWS     = Int32(8)    # the reference raises on this input
const AH_ST_ITER_LIMIT     = Int32(16)   # Remedy-3 loop guard of the library (the reference would keep looping)
const AH_ST_INTERLACE      = Int32(32)   # computed value violates the interlacing bounds (the reference is silent)
const REDO_MASK = AH_ST_NEEDS_BIGFLOAT | AH_ST_POLE_RETURN | AH_ST_REF_THROWS   # redo with the retained Julia solver
const WARN_MASK = AH_ST_ITER_LIMIT | AH_ST_INTERLACE                            # kept, reported to the caller

# BEGIN AhResult (parsed by tests/test_abi_layout.py -- keep one field per line)
mutable struct AhResult
    lambda::Ptr{Float64}
    sind::Ptr{Int64}
    kb::Ptr{Float64}
    kz::Ptr{Float64}
    knu::Ptr{Float64}
    krho::Ptr{Float64}
    qout::Ptr{Int64}
    steps::Ptr{Int32}
    status::Ptr{Int32}
    U::Ptr{Float64}
    ldu::Int64
    U_mem::Int32
    V::Ptr{Float64}
    ldv::Int64
    V_mem::Int32
    rowmap_u::Ptr{Int64}
    rowmap_v::Ptr{Int64}
    rowscale_v::Ptr{Float64}
end
# END AhResult

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        rc = ccall((:ah_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), CTX, 0)
        rc == 0 || error("ah_create: ", unsafe_string(ccall((:ah_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    CTX[]
end
check(rc) = rc == 0 || error("libarrowhead_b200 error $rc: ",
                             unsafe_string(ccall((:ah_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[])))

# several GPUs from this one Julia thread: ENV["ARROWHEAD_B200_GPUS"] = "8" -> ah_multi_* (worker threads live inside
# the library; the k-range is split into contiguous blocks, one per device, columns stream into the same host U)
const MULTI = Ref{Ptr{Cvoid}}(C_NULL)
function multi()
    ndev = parse(Int, get(ENV, "ARROWHEAD_B200_GPUS", "1"))
    ndev <= 1 && return C_NULL
    if MULTI[] == C_NULL
        rc = ccall((:ah_multi_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), MULTI, C_NULL, ndev)
        rc == 0 || error("ah_multi_create: ", unsafe_string(ccall((:ah_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    MULTI[]
end

struct Outputs
    lam::Vector{Float64}; sind::Vector{Int64}; kb::Vector{Float64}; kz::Vector{Float64}; knu::Vector{Float64}
    krho::Vector{Float64}; qout::Vector{Int64}; steps::Vector{Int32}; status::Vector{Int32}
end
Outputs(n) = Outputs(Vector{Float64}(undef, n), Vector{Int64}(undef, n), Vector{Float64}(undef, n), Vector{Float64}(undef, n),
                     Vector{Float64}(undef, n), Vector{Float64}(undef, n), Vector{Int64}(undef, n), Vector{Int32}(undef, n),
                     Vector{Int32}(undef, n))
result(o::Outputs, U, ldu, V, ldv, scale) =
    AhResult(pointer(o.lam), pointer(o.sind), pointer(o.kb), pointer(o.kz), pointer(o.knu), pointer(o.krho),
             pointer(o.qout), pointer(o.steps), pointer(o.status), U, ldu, AH_MEM_HOST, V, ldv, AH_MEM_HOST,
             C_NULL, C_NULL, scale)

# what the library could not finish the way the reference would: surfaced, never silently recomputed
function report(o::Outputs, what)
    w = findall(s -> s & WARN_MASK != 0, o.status)
    isempty(w) || @warn "$what: eigenpairs $w violate the interlacing bounds or hit the Remedy-3 loop guard " *
                        "(status & 0x30); the reference returns such values silently -- redo them in higher precision"
    return findall(s -> s & REDO_MASK != 0, o.status)
end

"""
    solve!(Λ, U, κ, Ax::SymArrow, zx)

Drop-in for `Threads.@threads for k=1:n; Λ[zx[k]],U[zx,zx[k]],κ[zx[k]] = eigen(Ax,Bx[tid],k); end`
(arrowhead3.jl:627-632 and 639-644).  `Ax` is ordered and irreducible with the arrow top last; `zx` maps
the reduced problem's rows/columns into the identity-initialised `U` (pass `1:n` when nothing was deflated).
τ is deliberately NOT forwarded: the reference's loop calls `eigen(Ax,Bx[tid],k)` with the per-k default.
"""
function solve!(Λ::Vector{Float64}, U::Matrix{Float64}, κ::Vector{Info}, Ax::SymArrow{Float64}, zx::AbstractVector{<:Integer})
    N = length(Ax.D); n = N + 1
    o = Outputs(n)
    deflated = length(zx) != size(U, 1) || zx != 1:n
    Ublk = deflated ? Matrix{Float64}(undef, n, n) : U            # contiguous n x n block for the reduced problem
    m = multi()
    GC.@preserve o Ublk begin
        r = result(o, pointer(Ublk), n, C_NULL, 0, C_NULL)
        if m == C_NULL
            check(ccall((:ah_symarrow_eigen, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Int64, Int64, Ref{AhResult}),
                        ctx(), N, Ax.D, Ax.z, Ax.a, C_NULL, 1, n, r))
        else
            check(ccall((:ah_multi_symarrow_eigen, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ref{AhResult}, Ptr{Ptr{Float64}}),
                        m, N, Ax.D, Ax.z, Ax.a, C_NULL, r, C_NULL))
        end
    end
    redo = report(o, "eigen(::SymArrow)")
    if !isempty(redo)                     # the retained per-k solver, arrowhead3.jl:419-550 (BigFloat / raising branches)
        B = deepcopy(Ax)
        for k in redo
            λ, v, info = parentmodule(@__MODULE__).eigen(Ax, B, k)
            o.lam[k] = λ; Ublk[:, k] = v
            o.sind[k] = info.Sind; o.kb[k] = info.Kb; o.kz[k] = info.Kz; o.knu[k] = info.Kν; o.krho[k] = info.Kρ; o.qout[k] = info.Qout
        end
    end
    deflated && (U[zx, zx] = Ublk)                               # U[zx,zx[k]] = v for all k
    for k = 1:n
        Λ[zx[k]] = o.lam[k]
        κ[zx[k]] = Info(o.sind[k], o.kb[k], o.kz[k], o.knu[k], o.krho[k], o.qout[k])
    end
    return nothing
end

"""
    solve!(Λ, U, κ, Ax::SymDPR1, zx)

Drop-in for `Threads.@threads for k=1:n; Λ[zx[k]],U[zx,zx[k]],κ[zx[k]] = eigen(Ax,Bx[tid],k); end`
(arrowhead4.jl:516-522 and 529-535).  `Ax = SymDPR1(D,z,ρ)` is ordered and irreducible with ρ > 0 -- the driver has
already flipped the sign of the matrix when `A.r < 0` (`signr`, arrowhead4.jl:443-449) and flips Λ back afterwards
(:542); nothing of that changes.  The 1x1 quick return (:456-471) never reaches this call.
"""
function solve!(Λ::Vector{Float64}, U::Matrix{Float64}, κ::Vector{Info}, Ax::SymDPR1{Float64}, zx::AbstractVector{<:Integer})
    n = length(Ax.D)
    n >= 2 || error("solve!(::SymDPR1): the 1x1 case is closed-form in the driver (arrowhead4.jl:456-471)")
    Ax.r > 0 || error("solve!(::SymDPR1): ρ must be > 0 here (the driver negates the matrix first, arrowhead4.jl:443-449)")
    o = Outputs(n)
    deflated = length(zx) != size(U, 1) || zx != 1:n
    Ublk = deflated ? Matrix{Float64}(undef, n, n) : U
    m = multi()
    GC.@preserve o Ublk begin
        r = result(o, pointer(Ublk), n, C_NULL, 0, C_NULL)
        if m == C_NULL
            check(ccall((:ah_symdpr1_eigen, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Int64, Int64, Ref{AhResult}),
                        ctx(), n, Ax.D, Ax.u, Ax.r, C_NULL, 1, n, r))
        else
            check(ccall((:ah_multi_symdpr1_eigen, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ref{AhResult}, Ptr{Ptr{Float64}}),
                        m, n, Ax.D, Ax.u, Ax.r, C_NULL, r, C_NULL))
        end
    end
    redo = report(o, "eigen(::SymDPR1)")
    if !isempty(redo)                     # arrowhead4.jl:289-422
        B = SymArrow(Vector{Float64}(undef, n - 1), Vector{Float64}(undef, n - 1), 1.0, n)
        for k in redo
            λ, v, info = parentmodule(@__MODULE__).eigen(Ax, B, k)
            o.lam[k] = λ; Ublk[:, k] = v
            o.sind[k] = info.Sind; o.kb[k] = info.Kb; o.kz[k] = info.Kz; o.knu[k] = info.Kν; o.krho[k] = info.Kρ; o.qout[k] = info.Qout
        end
    end
    deflated && (U[zx, zx] = Ublk)
    for k = 1:n
        Λ[zx[k]] = o.lam[k]
        κ[zx[k]] = Info(o.sind[k], o.kb[k], o.kz[k], o.knu[k], o.krho[k], o.qout[k])
    end
    return nothing
end

"""
    solve!(Σ, U, V, κ, Ax::HalfArrow, zx, zxv, signD)

Drop-in for the (serial, in the reference) loop of arrowhead6.jl:635-642

    for k=1:n
        Σ[zx[k]],U[zx,zx[k]],V[zxv,zx[k]],κ[zx[k]] = svd(Ax,Bx[tid],k)
        V[zxv[1:end-1],zx[k]] = V[zxv[1:end-1],zx[k]].*signD[zxv[1:end-1]]
    end

and, with `signD = nothing`, for the loop after D-deflation (:622-628).  `Ax = HalfArrow(D,z)` is ordered, D > 0,
z nonzero, `length(z) == length(D)` (no tip) or `length(D)+1` (tip); left vectors have `length(z)` rows, right vectors
`length(D)+1`.  The sign row-scaling of :641 is applied by the library while it stores V (`rowscale_v`).
"""
function solve!(Σ::Vector{Float64}, U::Matrix{Float64}, V::Matrix{Float64}, κ::Vector{Info}, Ax::HalfArrow{Float64},
                zx::AbstractVector{<:Integer}, zxv::AbstractVector{<:Integer}, signD::Union{Nothing,Vector{Float64}})
    nd = length(Ax.D); nz = length(Ax.z)
    o = Outputs(nz)
    Ublk = Matrix{Float64}(undef, nz, nz)
    Vblk = Matrix{Float64}(undef, nd + 1, nz)
    scale = signD === nothing ? nothing : Float64[signD[zxv[l]] for l = 1:nd]      # signD[zxv[1:end-1]]
    m = multi()
    GC.@preserve o Ublk Vblk scale begin
        r = result(o, pointer(Ublk), nz, pointer(Vblk), nd + 1, scale === nothing ? Ptr{Float64}(C_NULL) : pointer(scale))
        if m == C_NULL
            check(ccall((:ah_halfarrow_svd, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ref{AhResult}),
                        ctx(), nd, nz, Ax.D, Ax.z, C_NULL, 1, nz, r))
        else
            check(ccall((:ah_multi_halfarrow_svd, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{AhResult}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}),
                        m, nd, nz, Ax.D, Ax.z, C_NULL, r, C_NULL, C_NULL))
        end
    end
    redo = report(o, "svd(::HalfArrow)")
    if !isempty(redo)                     # arrowhead6.jl:344-508
        B = SymArrow(Vector{Float64}(undef, nd), Vector{Float64}(undef, nd), 1.0, nd + 1)
        for k in redo
            σ, u, v, info = parentmodule(@__MODULE__).svd(Ax, B, k)
            o.lam[k] = σ; Ublk[:, k] = u
            Vblk[:, k] = scale === nothing ? v : vcat(v[1:nd] .* scale, v[nd+1])
            o.sind[k] = info.Sind; o.kb[k] = info.Kb; o.kz[k] = info.Kz; o.knu[k] = info.Kν; o.krho[k] = info.Kρ; o.qout[k] = info.Qout
        end
    end
    U[zx, zx[1:nz]] = Ublk                                       # U[zx,zx[k]] = u, V[zxv,zx[k]] = v for all k
    V[zxv, zx[1:nz]] = Vblk
    for k = 1:nz
        Σ[zx[k]] = o.lam[k]
        κ[zx[k]] = Info(o.sind[k], o.kb[k], o.kz[k], o.knu[k], o.krho[k], o.qout[k])
    end
    return nothing
end

end # module
