# ApproxFunB200.jl -- the binding a maintainer adds on the reference side.
#
# UNTESTED HERE: neither this container nor the GPU box has Julia (SURVEY.md section 0, item 4).  The
# same C ABI is exercised through ctypes by tests/ and bench.py; this file shows the `ccall` stubs and
# the method overrides that make `Fun(f, S, n)`, `values`, `coefficients`, `f(x)` and `sample` run on
# libapproxfun_b200.so without touching user code.  No CUDA.jl, no array dispatch: plain `ccall`.
module ApproxFunB200

using ApproxFun
using ApproxFun: Chebyshev, Fourier, Laurent, Fun, Space, coefficients, space, domain
import ApproxFun.ApproxFunBase: TransformPlan, ITransformPlan, plan_transform, plan_itransform,
                                ClenshawPlan, clenshaw, fromcanonical
import ApproxFun: chebbisectioninv, chebnormalizedcumsum!, bisectioninv
import Base: *

const LIB = get(ENV, "APPROXFUN_B200_LIB", "libapproxfun_b200.so")

const CHEB1_FWD, CHEB1_INV, FOURIER_FWD, FOURIER_INV, LAURENT_FWD, LAURENT_INV = Cint.(0:5)

function check(st::Int32)
    st == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:afb_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, length(buf))
    msg = unsafe_string(pointer(buf))
    st == 2 ? throw(DimensionMismatch(msg)) : error("libapproxfun_b200: $msg")   # AFB_BAD_SIZE == 2
end

__init__() = check(ccall((:afb_init, LIB), Int32, (Int32,), -1))

# ---- plan objects: what FastTransforms.ChebyshevTransformPlan / FFTW.r2rFFTWPlan are upstream -------
mutable struct B200Plan{T}
    handle::Ptr{Cvoid}
    n::Int
    batch::Int
    function B200Plan{T}(kind::Cint, n::Integer, batch::Integer) where T
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:afb_plan_create, LIB), Int32,
                    (Ref{Ptr{Cvoid}}, Int32, Int64, Int64, Int64, Int64, UInt32),
                    h, kind, n, 0, batch, n, 0))
        p = new{T}(h[], n, batch)
        finalizer(q -> ccall((:afb_plan_destroy, LIB), Int32, (Ptr{Cvoid},), q.handle), p)
        p
    end
end

function *(p::B200Plan{T}, v::StridedVecOrMat{T}) where T
    size(v, 1) == p.n && size(v, 2) == p.batch || throw(DimensionMismatch("plan built for $(p.n) x $(p.batch)"))
    out = similar(v)
    GC.@preserve v out check(ccall((:afb_plan_execute_host, LIB), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}),
                                   p.handle, v, out))
    out
end

# ---- the reference's own entry points, overridden for Float64 / ComplexF64 -------------------------
for (SP, FWD, INV, T) in ((:Chebyshev, :CHEB1_FWD, :CHEB1_INV, :Float64),
                          (:Fourier, :FOURIER_FWD, :FOURIER_INV, :Float64),
                          (:Laurent, :LAURENT_FWD, :LAURENT_INV, :ComplexF64))
    @eval begin
        function plan_transform(sp::$SP, v::StridedVecOrMat{$T})
            pl = B200Plan{$T}($FWD, size(v, 1), size(v, 2))
            TransformPlan{$T,typeof(sp),false,typeof(pl)}(sp, pl)
        end
        function plan_itransform(sp::$SP, v::StridedVecOrMat{$T})
            pl = B200Plan{$T}($INV, size(v, 1), size(v, 2))
            ITransformPlan{$T,typeof(sp),false,typeof(pl)}(sp, pl)
        end
        *(P::TransformPlan{$T,<:$SP,false,<:B200Plan}, v::StridedVecOrMat{$T}) = P.plan * v
        *(P::ITransformPlan{$T,<:$SP,false,<:B200Plan}, v::StridedVecOrMat{$T}) = P.plan * v
    end
end

# points(::Chebyshev, n): descending first-kind grid (NEWS.md:92)
function ApproxFun.points(sp::Chebyshev{<:ApproxFun.IntervalOrSegment{Float64}}, n::Integer)
    d = domain(sp)
    out = Vector{Float64}(undef, n)
    check(ccall((:afb_chebpoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64),
                out, n, ApproxFun.leftendpoint(d), ApproxFun.rightendpoint(d)))
    out
end

# clenshaw(c, x, plan): vector and matrix-coefficient forms (src/Extras/sample.jl:56,68,80,94)
function clenshaw(c::Vector{Float64}, x::Vector{Float64}, ::ClenshawPlan{<:Chebyshev,Float64})
    y = similar(x)
    check(ccall((:afb_clenshaw_cheb, LIB), Int32, (Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                c, length(c), x, y, length(x)))
    y
end
function clenshaw(c::Matrix{Float64}, x::Vector{Float64}, ::ClenshawPlan{<:Chebyshev,Float64})
    size(c, 2) == length(x) || throw(DimensionMismatch("size(c,2) != length(x)"))
    y = similar(x)
    check(ccall((:afb_clenshaw_cheb_cols, LIB), Int32,
                (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
                c, size(c, 1), stride(c, 2), x, y, length(x)))
    y
end

# chebbisectioninv (src/Extras/sample.jl:55-101): the 47 fused Clenshaw bisections
function chebbisectioninv(c::AbstractVector{Float64}, xl::AbstractVector{Float64}, ::ClenshawPlan{<:Chebyshev,Float64})
    cc, u = Vector{Float64}(c), Vector{Float64}(xl)
    out = similar(u)
    check(ccall((:afb_cheb_bisect, LIB), Int32, (Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                cc, length(cc), u, out, length(u), 47))
    out
end
function chebbisectioninv(c::AbstractMatrix{Float64}, xl::AbstractVector{Float64}, ::ClenshawPlan{<:Chebyshev,Float64})
    @assert size(c)[2] == length(xl)
    cc, u = Matrix{Float64}(c), Vector{Float64}(xl)
    out = similar(u)
    check(ccall((:afb_cheb_bisect_cols, LIB), Int32,
                (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                cc, size(cc, 1), size(cc, 1), u, out, length(u), 47))
    out
end

# chebnormalizedcumsum! (src/Extras/sample.jl:170-175)
function chebnormalizedcumsum!(f::Vector{Float64})
    N = length(f)
    resize!(f, N + 1)
    check(ccall((:afb_cheb_normcumsum, LIB), Int32, (Ptr{Float64}, Int64, Int64, Int64, Int32), f, N, 1, N + 1, 1))
    f
end
function chebnormalizedcumsum!(f::Matrix{Float64})
    check(ccall((:afb_cheb_normcumsum, LIB), Int32, (Ptr{Float64}, Int64, Int64, Int64, Int32),
                f, size(f, 1), size(f, 2), size(f, 1), 0))
    f
end

# bisectioninv for Chebyshev Funs (src/Extras/sample.jl:103-109) with fromcanonical fused in the kernel
function bisectioninv(cf::Fun{<:Chebyshev,Float64}, x::Vector{Float64})
    d = domain(cf)
    out = similar(x)
    c = coefficients(cf)
    check(ccall((:afb_samplecdf_cheb, LIB), Int32,
                (Ptr{Float64}, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Int64),
                c, length(c), ApproxFun.leftendpoint(d), ApproxFun.rightendpoint(d), x, out, length(x)))
    out
end

# sample(f, n) with the uniforms drawn on the device (Philox4x32-10 counter i for sample i): no rand(n), no H2D traffic
function sample_b200(f::Fun{<:Chebyshev,Float64}, n::Integer; seed::Integer=0)
    cf = ApproxFun.normalizedcumsum(f)
    d = domain(f)
    c = coefficients(cf)
    out = Vector{Float64}(undef, n)
    check(ccall((:afb_sample_cheb, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, UInt64, Ptr{Float64}, Int64),
                c, length(c), ApproxFun.leftendpoint(d), ApproxFun.rightendpoint(d), seed, out, n))
    out
end

# ---- 2-D: transform!(::TensorSpace{Chebyshev,Chebyshev}, M) and the fused conditional sampler ------------------
function cheb2d_transform!(M::Matrix{Float64}; inverse::Bool=false)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:afb_plan_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Int64, Int64, Int64, Int64, UInt32),
                h, inverse ? 7 : 6, size(M, 1), size(M, 2), 1, size(M, 1), 0))
    try
        GC.@preserve M check(ccall((:afb_plan_execute_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], M, M))
    finally
        ccall((:afb_plan_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
    end
    M
end

# f(x_p, y_p) for a Chebyshev x Chebyshev coefficient matrix (canonical variables)
function clenshaw2d(C::Matrix{Float64}, x::Vector{Float64}, y::Vector{Float64})
    length(x) == length(y) || throw(DimensionMismatch("x and y differ in length"))
    out = similar(x)
    check(ccall((:afb_clenshaw_cheb2d, LIB), Int32,
                (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                C, size(C, 1), size(C, 2), x, y, out, length(x)))
    out
end

# lines 208-213 of src/Extras/sample.jl in one call:  fA = evaluate.(f.A, ry'); AB = CB*fA; chebnormalizedcumsum!(AB);
# rx = chebbisectioninv(AB, u); fromcanonical.(domain(f,1), rx)
function sample_conditional(A::Matrix{Float64}, CB::Matrix{Float64}, dA, d1, ry::Vector{Float64}, u::Vector{Float64})
    size(A, 2) == size(CB, 2) || throw(DimensionMismatch("rank mismatch"))
    rx = similar(u)
    check(ccall((:afb_sample2d_cheb_lowrank, LIB), Int32,
                (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Float64, Float64, Float64, Float64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                A, size(A, 1), CB, size(CB, 1), size(A, 2), ApproxFun.leftendpoint(dA), ApproxFun.rightendpoint(dA),
                ApproxFun.leftendpoint(d1), ApproxFun.rightendpoint(d1), ry, u, rx, length(u)))
    rx
end

# sibling plans on the same engine: kinds 8/9 = FastTransforms Val(2) (REDFT00), 10/11 = SinSpace, 12..15 = Taylor/Hardy
chebyshevtransform2(v::VecOrMat{Float64}) = B200Plan{Float64}(Cint(8), size(v, 1), size(v, 2)) * v
ichebyshevtransform2(c::VecOrMat{Float64}) = B200Plan{Float64}(Cint(9), size(c, 1), size(c, 2)) * c

# Padua: points(Chebyshev()^2, N) and the vector transform of Fun(f, Chebyshev()^2, N)  (kinds 16 / 17)
function paduapoints_b200(N::Integer, dx, dy)
    np = Ref{Int64}(0)
    check(ccall((:afb_paduapoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Ref{Int64}),
                C_NULL, N, ApproxFun.leftendpoint(dx), ApproxFun.rightendpoint(dx), ApproxFun.leftendpoint(dy),
                ApproxFun.rightendpoint(dy), np))
    xy = Matrix{Float64}(undef, np[], 2)
    check(ccall((:afb_paduapoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Ref{Int64}),
                xy, N, ApproxFun.leftendpoint(dx), ApproxFun.rightendpoint(dx), ApproxFun.leftendpoint(dy),
                ApproxFun.rightendpoint(dy), np))
    xy
end
paduatransform_b200(v::Vector{Float64}) = B200Plan{Float64}(Cint(16), length(v), 1) * v
ipaduatransform_b200(c::Vector{Float64}) = B200Plan{Float64}(Cint(17), length(c), 1) * c

end # module
