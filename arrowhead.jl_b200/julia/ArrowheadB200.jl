# ArrowheadB200.jl -- the reference-side binding of libarrowhead_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` (see DESIGN.md section 1).  It is the
# file a maintainer of ivanslapnicar/Arrowhead.jl would `include` after `src/arrowhead6.jl`; it replaces
# exactly the `Threads.@threads for k` loop bodies of the three drivers and leaves every other line
# (sorting, deflation, givens, the final views) untouched.  The Python mirror of the same calls
# (`arrowhead.jl_b200/drivers.py`) is what the test-suite exercises.
#
# struct layout must match `ah_result` in include/arrowhead_b200.h field by field.
module ArrowheadB200

using LinearAlgebra
import ..Arrowhead: SymArrow, SymDPR1, HalfArrow, Info

const LIB = get(ENV, "ARROWHEAD_B200_LIB", "libarrowhead_b200.so")
const AH_MEM_HOST   = Int32(0)
const AH_MEM_DEVICE = Int32(1)

mutable struct AhResult
    lambda::Ptr{Float64}; sind::Ptr{Int64}
    kb::Ptr{Float64}; kz::Ptr{Float64}; knu::Ptr{Float64}; krho::Ptr{Float64}
    qout::Ptr{Int64}; steps::Ptr{Int32}; status::Ptr{Int32}
    U::Ptr{Float64}; ldu::Int64; U_mem::Int32
    V::Ptr{Float64}; ldv::Int64; V_mem::Int32
    rowmap_u::Ptr{Int64}; rowmap_v::Ptr{Int64}; rowscale_v::Ptr{Float64}
end

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

"""
    solve!(Λ, U, κ, Ax::SymArrow, zx)

Drop-in for the loop `Threads.@threads for k=1:n; Λ[zx[k]],U[zx,zx[k]],κ[zx[k]] = eigen(Ax,Bx[tid],k); end`
(arrowhead3.jl:627-632 and 639-644).  `Ax` is ordered and irreducible with the arrow top last; `zx` maps
the reduced problem's rows/columns into the identity-initialised `U` (pass `1:n` when nothing was deflated).
τ is deliberately NOT forwarded: the reference's loop calls `eigen(Ax,Bx[tid],k)` with the per-k default.
"""
function solve!(Λ::Vector{Float64}, U::Matrix{Float64}, κ::Vector{Info}, Ax::SymArrow{Float64}, zx::AbstractVector{<:Integer})
    N = length(Ax.D); n = N + 1
    lam = Vector{Float64}(undef, n); sind = Vector{Int64}(undef, n); qout = Vector{Int64}(undef, n)
    kb = similar(lam); kz = similar(lam); knu = similar(lam); krho = similar(lam)
    steps = Vector{Int32}(undef, n); status = Vector{Int32}(undef, n)
    deflated = length(zx) != size(U, 1) || zx != 1:n
    Ublk = deflated ? Matrix{Float64}(undef, n, n) : U            # contiguous n x n block for the reduced problem
    GC.@preserve lam sind kb kz knu krho qout steps status Ublk begin
        r = AhResult(pointer(lam), pointer(sind), pointer(kb), pointer(kz), pointer(knu), pointer(krho),
                     pointer(qout), pointer(steps), pointer(status),
                     pointer(Ublk), n, AH_MEM_HOST, C_NULL, 0, AH_MEM_HOST, C_NULL, C_NULL, C_NULL)
        check(ccall((:ah_symarrow_eigen, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Int64, Int64, Ref{AhResult}),
                    ctx(), N, Ax.D, Ax.z, Ax.a, C_NULL, 1, n, r))
    end
    any(s -> s & 0x2e != 0, status) && redo_on_host!(lam, Ublk, sind, kb, kz, knu, krho, qout, status, Ax)
    deflated && (U[zx, zx] = Ublk)                               # U[zx,zx[k]] = v for all k
    for k = 1:n
        Λ[zx[k]] = lam[k]
        κ[zx[k]] = Info(sind[k], kb[k], kz[k], knu[k], krho[k], qout[k])
    end
    return nothing
end

# Eigenpairs the GPU flags as "reference switches to BigFloat / raises here" go through the retained
# Julia per-k solver (arrowhead3.jl:419-550), exactly as before.
function redo_on_host!(lam, Ublk, sind, kb, kz, knu, krho, qout, status, Ax)
    B = deepcopy(Ax)
    for k in findall(s -> s & 0x2e != 0, status)
        λ, v, info = Main.Arrowhead.eigen(Ax, B, k)
        lam[k] = λ; Ublk[:, k] = v
        sind[k] = info.Sind; kb[k] = info.Kb; kz[k] = info.Kz; knu[k] = info.Kν; krho[k] = info.Kρ; qout[k] = info.Qout
    end
end

# SymDPR1 (arrowhead4.jl:516-522, 529-535) and HalfArrow (arrowhead6.jl:622-628, 635-642) follow the
# same pattern with ah_symdpr1_eigen (n rows per vector, rho > 0 after the `signr` flip at :443-449) and
# ah_halfarrow_svd (U: nz rows, V: nd+1 rows, rowscale_v = signD of arrowhead6.jl:641).

end # module
