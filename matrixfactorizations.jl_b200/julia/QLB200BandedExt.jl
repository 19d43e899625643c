# QLB200BandedExt.jl -- optional companion of QLB200.jl for BandedMatrices.jl users (SURVEY.md 8(f) N4).
# NOT EXECUTED IN THIS REPOSITORY (no `julia` binary in the image); the executable twin is ql.py:banded_ql_ (tests/test_banded.py).
#
# The reference's extension defines `ql!(A::BandedMatrix) = banded_ql!(A)` (ext/MatrixFactorizationsBandedMatricesExt.jl:16) after
# `ql(A::BandedMatrix)` has widened the lower bandwidth (ext:14).  A more specific method for Float64 routes the factorization of
# the band data to libqlb200 (one ccall); Q-applies and solves of the banded QL stay with the extension's loops (ext:38-185).
module QLB200BandedExt

using BandedMatrices, MatrixFactorizations, LinearAlgebra
using BandedMatrices: bandeddata
import MatrixFactorizations: ql!, QL
import ..QLB200: handle, chk, libqlb200

function ql!(L::BandedMatrix{Float64})                     # replaces banded_ql!(L) (ext:18-36)
    D = bandeddata(L)                                      # (l+u+1) x n, column-major
    l, u = bandwidths(L)
    m, n = size(L)
    τ = zeros(Float64, min(m, n))
    h = handle()
    info = ccall((:qlb200_dgbqlf, libqlb200), Cint,
                 (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                 h.ptr, m, n, l, u, D, max(1, stride(D, 2)), τ)
    chk(h, info)
    QL(L, τ)
end

end # module
