# QLB200.jl -- Julia-side shim that routes MatrixFactorizations.jl's QL hot path to libqlb200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary (SURVEY.md 0.7), so this file is
# the reference-side binding a maintainer would add (as a package extension or a tiny companion package); the
# executable mirror of the same calls is ../ql.py (ctypes) and is what tests/ drive.  Every method below is a
# single `ccall` into the C-ABI of include/qlb200.h; nothing here does arithmetic.
#
# Hooks overridden (all exist in the reference, precedent: ext/MatrixFactorizationsBandedMatricesExt.jl:14-16,38,66):
#   qlfactUnblocked!(::StridedMatrix{T}), T in Float64/Float32/ComplexF64/ComplexF32   src/ql.jl:72
#   materialize!(::Lmul{<:QLPackedQLayout{<:AbstractColumnMajor}})     src/ql.jl:321
#   materialize!(::Lmul{<:AdjQLPackedQLayout{<:AbstractColumnMajor}})  src/ql.jl:350
#   materialize!(::Rmul{<:Any,<:QLPackedQLayout}) / Adj                src/ql.jl:381,409
#   materialize!(::Ldiv{<:QLPackedLayout{<:AbstractColumnMajor}})      src/ql.jl:440 (square; Float64 also tall least squares)
#   rq!(::StridedMatrix{Float64})                                      src/rq.jl:401
#   lmul!/rmul!(::RQPackedQ{Float64,<:StridedMatrix}, ...)             src/rq.jl:130,183,234,272
#   ul!(::StridedMatrix{Float64}, pivot; check)                        src/ul.jl:146
# plus the new batched surface ql_batched!/lmul_batched!/ldiv_batched!/ql_apply_batched! (SURVEY.md 0.6; any n <= 64).
module QLB200

using LinearAlgebra, MatrixFactorizations
using MatrixFactorizations.ArrayLayouts
using LinearAlgebra: BlasInt
import MatrixFactorizations: qlfactUnblocked!, rq!, ul!, QL, QLPackedQ, RQ, RQPackedQ, UL,
                             QLPackedQLayout, AdjQLPackedQLayout, QLPackedLayout,
                             _qrfactUnblocked!, QR, QRPackedQ
import ArrayLayouts: materialize!, Lmul, Rmul, Ldiv, AbstractColumnMajor, AbstractStridedLayout,
                     QRPackedQLayout, AdjQRPackedQLayout

const libqlb200 = get(ENV, "QLB200_LIB", joinpath(@__DIR__, "..", "libqlb200.so"))

# ---- handle ---------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        info = ccall((:qlb200_create, libqlb200), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device)
        info == 0 || error("qlb200_create(device=$device) failed with info=$info (no CUDA device? there is no CPU fallback)")
        h = new(r[])
        finalizer(x -> ccall((:qlb200_destroy, libqlb200), Cint, (Ptr{Cvoid},), x.ptr), h)
        h
    end
end
const DEFAULT = Ref{Union{Nothing,Handle}}(nothing)
handle() = (DEFAULT[] === nothing && (DEFAULT[] = Handle(0)); DEFAULT[]::Handle)

lasterror(h::Handle) = unsafe_string(ccall((:qlb200_last_error, libqlb200), Cstring, (Ptr{Cvoid},), h.ptr))

# info -> the exception the reference throws at the same place
function chk(h::Handle, info::Integer; dims = ())
    info == 0 && return nothing
    msg = lasterror(h)
    info >= 1000 && error("libqlb200: CUDA error $(info - 1000): $msg")
    info > 0 && throw(SingularException(info))
    (-info in dims || occursin("DimensionMismatch", msg)) && throw(DimensionMismatch(msg))
    throw(ArgumentError("libqlb200: invalid argument #$(-info): $msg"))
end

# Page-lock a Julia array for the host-pointer entry points (cudaHostRegister inside the library): copies then run at full PCIe
# speed (about 5x the pageable rate) and a repeated large `ql!` on the same array replays its whole upload | factor | download
# schedule as one CUDA graph.  `unpin!` before the array can be garbage-collected.
pin!(A::StridedArray{<:LinearAlgebra.BlasFloat}; h::Handle = handle()) =
    (chk(h, ccall((:qlb200_host_register, libqlb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), h.ptr, pointer(A), sizeof(eltype(A)) * (stride(A, ndims(A)) * size(A, ndims(A))))); A)
unpin!(A::StridedArray{<:LinearAlgebra.BlasFloat}; h::Handle = handle()) =
    (chk(h, ccall((:qlb200_host_unregister, libqlb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, pointer(A))); A)

# ---- ql!, lmul!/rmul!, ldiv! for every BlasFloat element type ------------------------------------------------------
# replaces `QL(LAPACK.geqlf!(A)...)` (src/ql.jl:72); ql!/ql/ql_layout! reach it unchanged (src/ql.jl:114-117,179-191).
# Float64: native; ComplexF64: native complex panel + the same DMMA GEMMs (csrc/complex.cu); Float32 / ComplexF32: widened
# on the device (csrc/mixed.cu).  For complex types the adjoint is 'C' (?unmql), the conj placements of
# src/ql.jl:336,366,368,400,426,429 are inside the library.
for (T, p, orm, adj) in ((Float64, :d, :ormql, 'T'), (Float32, :s, :ormql, 'T'), (ComplexF64, :z, :unmql, 'C'), (ComplexF32, :c, :unmql, 'C'))
    geqlf = QuoteNode(Symbol(:qlb200_, p, :geqlf))
    ormql = QuoteNode(Symbol(:qlb200_, p, orm))
    solve = QuoteNode(Symbol(:qlb200_, p, :qlsolve))
    @eval begin
        function qlfactUnblocked!(A::StridedMatrix{$T})
            Base.require_one_based_indexing(A)
            LinearAlgebra.chkstride1(A)
            m, n = size(A)
            τ = Vector{$T}(undef, min(m, n))
            h = handle()
            info = ccall(($geqlf, libqlb200), Cint,
                         (Ptr{Cvoid}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}),
                         h.ptr, m, n, A, max(1, stride(A, 2)), τ)
            chk(h, info)
            QL(A, τ)
        end
        function _ormql!(side::Char, trans::Char, A::QLPackedQ{$T,<:StridedMatrix}, C::StridedVecOrMat{$T})
            Base.require_one_based_indexing(C)
            mA, nA = size(A.factors)
            k = min(mA, nA)
            m, n = size(C, 1), size(C, 2)
            nq = side == 'L' ? m : n
            nq == mA || throw(DimensionMismatch("matrix A has dimensions ($mA,$nA) but $(side == 'L' ? "B" : "A") has dimensions ($m, $n)"))
            V = view(A.factors, :, nA-k+1:nA)                  # the reflectors live in the last k columns (src/ql.jl:331)
            h = handle()
            info = ccall(($ormql, libqlb200), Cint,
                         (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Ptr{$T}, Int64),
                         h.ptr, side, trans == 'N' ? 'N' : $adj, m, n, k, V, max(1, stride(A.factors, 2)), A.τ, C, max(1, stride(C, 2)))
            chk(h, info; dims = (6, 8))
            C
        end
        materialize!(M::Lmul{<:QLPackedQLayout{<:AbstractColumnMajor},<:AbstractColumnMajor,<:QLPackedQ{$T},<:StridedVecOrMat{$T}}) =
            _ormql!('L', 'N', M.A, M.B)                                                           # src/ql.jl:321-347
        materialize!(M::Lmul{<:AdjQLPackedQLayout{<:AbstractColumnMajor},<:AbstractColumnMajor,<:Any,<:StridedVecOrMat{$T}}) =
            _ormql!('L', 'T', parent(M.A), M.B)                                                   # src/ql.jl:350-377
        materialize!(M::Rmul{<:AbstractColumnMajor,<:QLPackedQLayout{<:AbstractColumnMajor},<:StridedMatrix{$T},<:QLPackedQ{$T}}) =
            _ormql!('R', 'N', M.B, M.A)                                                           # src/ql.jl:381-406
        materialize!(M::Rmul{<:AbstractColumnMajor,<:AdjQLPackedQLayout{<:AbstractColumnMajor},<:StridedMatrix{$T}}) =
            _ormql!('R', 'T', parent(M.B), M.A)                                                   # src/ql.jl:409-435

        # ldiv!(F::QL, B): square (src/ql.jl:440-447,467); Float64 also the tall least-squares case (SURVEY.md 8(f) N2)
        function materialize!(Ldv::Ldiv{<:QLPackedLayout{<:AbstractColumnMajor},<:AbstractColumnMajor,<:QL{$T},<:StridedVecOrMat{$T}})
            F, B = Ldv.A, Ldv.B
            m, n = size(F)
            if m > n && $T === Float64      # the reference's own branch (src/ql.jl:448-486) is @test_broken
                size(B, 1) == m || throw(DimensionMismatch("arguments must have the same number of rows; has $m and $(size(B,1))"))
                h = handle()
                info = ccall((:qlb200_dqllstsq, libqlb200), Cint,
                             (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                             h.ptr, m, n, size(B, 2), F.factors, max(1, stride(F.factors, 2)), F.τ, B, max(1, stride(B, 2)))
                chk(h, info; dims = (9,))
                return B      # X = B[1:n, :] (what `\` cuts out, src/ql.jl:501); B[n+1:m, :] = coordinates of the residual
            end
            if m < n && $T === Float64      # wide: minimum-norm solution on the n-row buffer `\` allocates (src/ql.jl:496-501)
                size(B, 1) == n || throw(DimensionMismatch("wide ldiv!: B must have $n rows (b in the first $m); has $(size(B,1))"))
                h = handle()
                info = ccall((:qlb200_dqlminnorm, libqlb200), Cint,
                             (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                             h.ptr, m, n, size(B, 2), F.factors, max(1, stride(F.factors, 2)), F.τ, B, max(1, stride(B, 2)))
                chk(h, info; dims = (9,))
                return B
            end
            m == n || return invoke(materialize!, Tuple{Ldiv{<:QLPackedLayout}}, Ldv)   # other element types, rectangular: the reference's own branch (src/ql.jl:448-486)
            size(B, 1) == n || throw(DimensionMismatch("arguments must have the same number of rows; has $n and $(size(B,1))"))
            h = handle()
            info = ccall(($solve, libqlb200), Cint,
                         (Ptr{Cvoid}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Ptr{$T}, Int64),
                         h.ptr, n, size(B, 2), F.factors, max(1, stride(F.factors, 2)), F.τ, B, max(1, stride(B, 2)))
            chk(h, info; dims = (8,))
            B
        end
    end
end

# ---- qrunblocked! and QR's packed Q through index reversal (SURVEY.md 8(f) N3) ------------------------------------
function _qrfactUnblocked!(::AbstractColumnMajor, ::AbstractStridedLayout, A::AbstractMatrix{Float64}, τ::AbstractVector{Float64})   # src/qr.jl:72
    m, n = size(A)
    length(τ) >= min(m, n) || throw(DimensionMismatch("tau has length $(length(τ)), needs $(min(m,n))"))
    h = handle()
    chk(h, ccall((:qlb200_dgeqrf, libqlb200), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                 h.ptr, m, n, A, max(1, stride(A, 2)), τ))
    QR(A, τ)
end
function _ormqr!(side::Char, trans::Char, Q::QRPackedQ{Float64,<:StridedMatrix}, C::StridedVecOrMat{Float64})
    mA, nA = size(Q.factors)
    k = min(mA, nA)
    m, n = size(C, 1), size(C, 2)
    (side == 'L' ? m : n) == mA || throw(DimensionMismatch("matrix A has dimensions ($mA,$nA) but B has dimensions ($m, $n)"))
    h = handle()
    chk(h, ccall((:qlb200_dormqr, libqlb200), Cint,
                 (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                 h.ptr, side, trans, m, n, k, Q.factors, max(1, stride(Q.factors, 2)), Q.τ, C, max(1, stride(C, 2))); dims = (6, 8))
    C
end
materialize!(M::Lmul{<:QRPackedQLayout{<:AbstractColumnMajor},<:AbstractColumnMajor,<:QRPackedQ{Float64},<:StridedVecOrMat{Float64}}) =
    _ormqr!('L', 'N', M.A, M.B)
materialize!(M::Lmul{<:AdjQRPackedQLayout{<:AbstractColumnMajor},<:AbstractColumnMajor,<:Any,<:StridedVecOrMat{Float64}}) =
    _ormqr!('L', 'T', parent(M.A), M.B)
materialize!(M::Rmul{<:AbstractColumnMajor,<:QRPackedQLayout{<:AbstractColumnMajor},<:StridedMatrix{Float64},<:QRPackedQ{Float64}}) =
    _ormqr!('R', 'N', M.B, M.A)
materialize!(M::Rmul{<:AbstractColumnMajor,<:AdjQRPackedQLayout{<:AbstractColumnMajor},<:StridedMatrix{Float64}}) =
    _ormqr!('R', 'T', parent(M.B), M.A)

# ---- rq! and RQ's packed Q (every BlasFloat element type) ---------------------------------------------------------------
for (T, p, orm, adj) in ((Float64, :d, :ormrq, 'T'), (Float32, :s, :ormrq, 'T'), (ComplexF64, :z, :unmrq, 'C'), (ComplexF32, :c, :unmrq, 'C'))
    gerqf = QuoteNode(Symbol(:qlb200_, p, :gerqf))
    ormrq = QuoteNode(Symbol(:qlb200_, p, orm))
    @eval begin
        function rq!(A::StridedMatrix{$T})                                                    # src/rq.jl:401
            m, n = size(A)
            τ = Vector{$T}(undef, min(m, n))
            h = handle()
            chk(h, ccall(($gerqf, libqlb200), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}),
                         h.ptr, m, n, A, max(1, stride(A, 2)), τ))
            RQ(A, τ)
        end
        function _ormrq!(side::Char, trans::Char, Q::RQPackedQ{$T,<:StridedMatrix}, C::StridedVecOrMat{$T})
            mA, nA = size(Q.factors)
            k = min(mA, nA)
            m, n = size(C, 1), size(C, 2)
            (side == 'L' ? m : n) == nA || throw(DimensionMismatch("matrix A has dimensions ($mA,$nA) but B has dimensions ($m, $n)"))
            V = view(Q.factors, mA-k+1:mA, :)                  # the last k rows hold the reflectors (LAPACK ?ormrq / ?unmrq)
            h = handle()
            chk(h, ccall(($ormrq, libqlb200), Cint,
                         (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Ptr{$T}, Int64),
                         h.ptr, side, trans == 'N' ? 'N' : $adj, m, n, k, V, max(1, stride(Q.factors, 2)), Q.τ, C, max(1, stride(C, 2))); dims = (6, 8))
            C
        end
        LinearAlgebra.lmul!(Q::RQPackedQ{$T,<:StridedMatrix}, B::StridedVecOrMat{$T}) = _ormrq!('L', 'N', Q, B)          # src/rq.jl:130-137
        LinearAlgebra.lmul!(Q::Adjoint{<:Any,<:RQPackedQ{$T,<:StridedMatrix}}, B::StridedVecOrMat{$T}) = _ormrq!('L', 'T', parent(Q), B)   # :183-202
        LinearAlgebra.rmul!(A::StridedMatrix{$T}, Q::RQPackedQ{$T,<:StridedMatrix}) = _ormrq!('R', 'N', Q, A)           # :234-242
        LinearAlgebra.rmul!(A::StridedMatrix{$T}, Q::Adjoint{<:Any,<:RQPackedQ{$T,<:StridedMatrix}}) = _ormrq!('R', 'T', parent(Q), A)   # :272-291
    end
end

# ---- ul! -----------------------------------------------------------------------------------------------------------
function ul!(A::StridedMatrix{Float64}, ::Val{Pivot} = Val(true); check::Bool = true) where Pivot        # src/ul.jl:146-196
    m, n = size(A)
    ipiv = Vector{BlasInt}(undef, min(m, n))          # BlasInt == Int64 on 64-bit Julia
    info = Ref{Int32}(0)
    h = handle()
    chk(h, ccall((:qlb200_dulfact, libqlb200), Cint,
                 (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Cint, Ref{Int32}),
                 h.ptr, m, n, A, max(1, stride(A, 2)), ipiv, Pivot ? 1 : 0, info))
    check && MatrixFactorizations._checknonsingular(info[], Val{Pivot}())                                 # src/ul.jl:193
    UL{Float64}(A, ipiv, convert(BlasInt, info[]))
end

# ---- batched surface (new; per-matrix semantics = ql!, lmul!, ldiv!) --------------------------------------------------
for (T, p) in ((Float64, :d), (Float32, :s))
    geqlf = QuoteNode(Symbol(:qlb200_, p, :geqlf_batched))
    ormql = QuoteNode(Symbol(:qlb200_, p, :ormql_batched))
    solve = QuoteNode(Symbol(:qlb200_, p, :qlsolve_batched))
    fused = QuoteNode(Symbol(:qlb200_, p, :geqlf_apply_batched))
    @eval begin
        """`ql_batched!(A::Array{T,3})`: `[ql!(view(A,:,:,i)) for i]` in one call; returns the vector of `QL`s."""
        function ql_batched!(A::Array{$T,3})
            n, n2, b = size(A)
            n == n2 || throw(DimensionMismatch("ql_batched!: square matrices only"))
            τ = Matrix{$T}(undef, n, b)
            h = handle()
            chk(h, ccall(($geqlf, libqlb200), Cint, (Ptr{Cvoid}, Int64, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Int64),
                         h.ptr, n, A, n, n * n, τ, n, b))
            [QL(view(A, :, :, i), view(τ, :, i)) for i in 1:b]
        end
        function lmul_batched!(trans::Char, A::Array{$T,3}, τ::Matrix{$T}, B::Array{$T,3})
            n, _, b = size(A)
            size(B, 1) == n && size(B, 3) == b || throw(DimensionMismatch("lmul_batched!"))
            h = handle()
            chk(h, ccall(($ormql, libqlb200), Cint,
                         (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64),
                         h.ptr, trans, n, size(B, 2), A, n, n * n, τ, n, B, n, n * size(B, 2), b))
            B
        end
        """`ql_apply_batched!(op, A, B)`: `F_i = ql!(view(A,:,:,i))` followed by `lmul!(F_i.Q', B_i)` (op 'T'), `lmul!(F_i.Q, B_i)`
        ('N') or `ldiv!(F_i, B_i)` ('S') in ONE call: the factors cross PCIe once.  Returns (A, τ, B)."""
        function ql_apply_batched!(op::Char, A::Array{$T,3}, B::Array{$T,3})
            n, n2, b = size(A)
            n == n2 && size(B, 1) == n && size(B, 3) == b || throw(DimensionMismatch("ql_apply_batched!"))
            τ = Matrix{$T}(undef, n, b)
            info = Vector{Int32}(undef, b)
            h = handle()
            chk(h, ccall(($fused, libqlb200), Cint,
                         (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64, Ptr{Int32}),
                         h.ptr, op, n, size(B, 2), A, n, n * n, τ, n, B, n, n * size(B, 2), b, info))
            if op == 'S'
                i = findfirst(!iszero, info)
                i === nothing || throw(SingularException(info[i]))
            end
            A, τ, B
        end
        function ldiv_batched!(A::Array{$T,3}, τ::Matrix{$T}, B::Array{$T,3})
            n, _, b = size(A)
            size(B, 1) == n && size(B, 3) == b || throw(DimensionMismatch("ldiv_batched!"))
            info = Vector{Int32}(undef, b)
            h = handle()
            chk(h, ccall(($solve, libqlb200), Cint,
                         (Ptr{Cvoid}, Int64, Int64, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64, Ptr{Int32}),
                         h.ptr, n, size(B, 2), A, n, n * n, τ, n, B, n, n * size(B, 2), b, info))
            i = findfirst(!iszero, info)
            i === nothing || throw(SingularException(info[i]))
            B
        end
    end
end

end # module
