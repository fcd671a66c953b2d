# ApproxFunB200.jl -- the binding a maintainer adds on the reference side.
#
# UNTESTED HERE: neither this container nor the GPU box has Julia (SURVEY.md section 0, item 4).  The
# same C ABI is exercised through ctypes by tests/ and bench.py and by a plain-C caller (tests/c/abi_smoke.c);
# this file shows the `ccall` stubs and the METHOD OVERRIDES that make `Fun(f, S, n)`, `values`, `coefficients`,
# `f(x)`, `transform!(::TensorSpace, M)` and `sample` run on libapproxfun_b200.so without touching user code.
# No CUDA.jl, no array dispatch: plain `ccall` on the arrays Julia already owns.
module ApproxFunB200

using ApproxFun
using ApproxFun: Chebyshev, Fourier, Laurent, Jacobi, Ultraspherical, Fun, Space, TensorSpace, coefficients, space, domain,
                 leftendpoint, rightendpoint, ncoefficients
import ApproxFun.ApproxFunBase: TransformPlan, ITransformPlan, plan_transform, plan_itransform, plan_transform!,
                                plan_itransform!, transform!, itransform!, ClenshawPlan, clenshaw, fromcanonical, tocanonical
import ApproxFun: chebbisectioninv, chebnormalizedcumsum!, bisectioninv
import ApproxFun.ApproxFunOrthogonalPolynomials: recA, recB, recC
import Base: *

const LIB = get(ENV, "APPROXFUN_B200_LIB", "libapproxfun_b200.so")

const CHEB1_FWD, CHEB1_INV, FOURIER_FWD, FOURIER_INV, LAURENT_FWD, LAURENT_INV, CHEB1_2D_FWD, CHEB1_2D_INV = Cint.(0:7)
const FOURIER_CHEB1_2D_FWD, FOURIER_CHEB1_2D_INV, LAURENT_CHEB1_2D_FWD, LAURENT_CHEB1_2D_INV = Cint.(18:21)

function check(st::Int32)
    st == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:afb_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, length(buf))
    msg = unsafe_string(pointer(buf))
    st == 2 ? throw(DimensionMismatch(msg)) : error("libapproxfun_b200: $msg")   # AFB_BAD_SIZE == 2
end

# afb_init(devices, ndev): ndev == 0 keeps the current device.  One Julia process can drive every GPU of the node:
#     ApproxFunB200.use_devices(0:7)
# after which plan*v on a Matrix shards its columns over the GPUs, transform!(::TensorSpace, M) runs as column slabs ->
# NVLink peer-memory transpose -> row slabs, and clenshaw / sample shard their points (all inside the one ccall).
function use_devices(devs = Int32[])
    d = collect(Int32, devs)
    check(ccall((:afb_init, LIB), Int32, (Ptr{Int32}, Int32), d, length(d)))
end
__init__() = use_devices()

# Page-lock a work array in place so host calls DMA it directly (a pageable Array is staged through the library's pinned
# ring at about 0.4x the speed).  The array must be unpinned before the GC frees it: pin!(A) installs that finalizer.
function pin!(A::Array)
    check(ccall((:afb_host_register, LIB), Int32, (Ptr{Cvoid}, Csize_t), A, sizeof(A)))
    finalizer(a -> ccall((:afb_host_unregister, LIB), Int32, (Ptr{Cvoid},), a), A)
    A
end
unpin!(A::Array) = check(ccall((:afb_host_unregister, LIB), Int32, (Ptr{Cvoid},), A))

# ---- plan objects: what FastTransforms.ChebyshevTransformPlan / FFTW.r2rFFTWPlan are upstream -------
mutable struct B200Plan{T}
    handle::Ptr{Cvoid}
    kind::Cint
    n::Int
    batch::Int
    stride::Int
    function B200Plan{T}(kind::Cint, n::Integer, batch::Integer, stride::Integer = n; n2::Integer = 0) where T
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:afb_plan_create, LIB), Int32,
                    (Ref{Ptr{Cvoid}}, Int32, Int64, Int64, Int64, Int64, UInt32),
                    h, kind, n, n2, batch, stride, 0))
        p = new{T}(h[], kind, n, n2 == 0 ? batch : n2, stride)
        finalizer(q -> ccall((:afb_plan_destroy, LIB), Int32, (Ptr{Cvoid},), q.handle), p)
        p
    end
end

# The library reads `batch` lines of length n, `stride` elements apart, from a raw pointer: only arrays whose memory layout
# is exactly that may be passed.  Dense arrays qualify (stride = n); so does a column range of a larger matrix
# (unit stride along dimension 1, stride(v, 2) between columns) -- the plan is then created with that stride.  Anything else
# (non-unit stride(v, 1), reshaped views ...) is copied first.
layout_ok(v::Array) = true
layout_ok(v::StridedVecOrMat) = stride(v, 1) == 1 && (ndims(v) == 1 || size(v, 2) == 1 || stride(v, 2) >= size(v, 1))
linestride(v::StridedVecOrMat) = (ndims(v) == 1 || size(v, 2) == 1) ? size(v, 1) : stride(v, 2)

function *(p::B200Plan{T}, v::StridedVecOrMat{T}) where T
    size(v, 1) == p.n && size(v, 2) == p.batch || throw(DimensionMismatch("plan built for $(p.n) x $(p.batch)"))
    if !(layout_ok(v) && linestride(v) == p.stride)                        # any other layout: a dense copy through a dense plan
        return B200Plan{T}(p.kind, p.n, p.batch) * copy(v)
    end
    out = similar(v)
    GC.@preserve v out check(ccall((:afb_plan_execute_host, LIB), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}),
                                   p.handle, pointer(v), pointer(out)))
    out
end
# in place on the caller's (possibly strided) memory: the library allows in == out
function execute!(p::B200Plan{T}, v::StridedVecOrMat{T}) where T
    layout_ok(v) && linestride(v) == p.stride || throw(ArgumentError("plan stride $(p.stride) does not match the array"))
    GC.@preserve v check(ccall((:afb_plan_execute_host, LIB), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}), p.handle, pointer(v), pointer(v)))
    v
end

# ---- the reference's own entry points, overridden for Float64 / ComplexF64 -------------------------
for (SP, FWD, INV, T) in ((:Chebyshev, :CHEB1_FWD, :CHEB1_INV, :Float64),
                          (:Fourier, :FOURIER_FWD, :FOURIER_INV, :Float64),
                          (:Laurent, :LAURENT_FWD, :LAURENT_INV, :ComplexF64))
    @eval begin
        function plan_transform(sp::$SP, v::StridedVecOrMat{$T})
            vv = layout_ok(v) ? v : copy(v)
            pl = B200Plan{$T}($FWD, size(vv, 1), size(vv, 2), linestride(vv))
            TransformPlan{$T,typeof(sp),false,typeof(pl)}(sp, pl)
        end
        function plan_itransform(sp::$SP, v::StridedVecOrMat{$T})
            vv = layout_ok(v) ? v : copy(v)
            pl = B200Plan{$T}($INV, size(vv, 1), size(vv, 2), linestride(vv))
            ITransformPlan{$T,typeof(sp),false,typeof(pl)}(sp, pl)
        end
        # plan_transform(sp, n::Integer) (NEWS.md:90; call sites src/Plot/Plot.jl:290-332): a plan for one vector of length n
        plan_transform(sp::$SP, n::Integer) = plan_transform(sp, Vector{$T}(undef, n))
        plan_itransform(sp::$SP, n::Integer) = plan_itransform(sp, Vector{$T}(undef, n))
        # in-place plans: same kernels with in == out
        function plan_transform!(sp::$SP, v::StridedVecOrMat{$T})
            pl = B200Plan{$T}($FWD, size(v, 1), size(v, 2), linestride(v))
            TransformPlan{$T,typeof(sp),true,typeof(pl)}(sp, pl)
        end
        function plan_itransform!(sp::$SP, v::StridedVecOrMat{$T})
            pl = B200Plan{$T}($INV, size(v, 1), size(v, 2), linestride(v))
            ITransformPlan{$T,typeof(sp),true,typeof(pl)}(sp, pl)
        end
        *(P::TransformPlan{$T,<:$SP,false,<:B200Plan}, v::StridedVecOrMat{$T}) = P.plan * v
        *(P::ITransformPlan{$T,<:$SP,false,<:B200Plan}, v::StridedVecOrMat{$T}) = P.plan * v
        *(P::TransformPlan{$T,<:$SP,true,<:B200Plan}, v::StridedVecOrMat{$T}) = execute!(P.plan, v)
        *(P::ITransformPlan{$T,<:$SP,true,<:B200Plan}, v::StridedVecOrMat{$T}) = execute!(P.plan, v)
    end
end

# points(::Chebyshev, n): descending first-kind grid (NEWS.md:92)
function ApproxFun.points(sp::Chebyshev{<:ApproxFun.IntervalOrSegment{Float64}}, n::Integer)
    d = domain(sp)
    out = Vector{Float64}(undef, n)
    check(ccall((:afb_chebpoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64),
                out, n, leftendpoint(d), rightendpoint(d)))
    out
end

# ---- 2-D: transform!(::TensorSpace, M) / itransform! (UPSTREAM ApproxFunBase; callers src/Plot/Plot.jl:290-332,
#      src/Extras/sample.jl:196, test/PeriodicXIntervalTest.jl:22-24): factor-1 plan on every column, factor-2 plan on every row
const ChebCheb = TensorSpace{<:Tuple{<:Chebyshev,<:Chebyshev}}
const FourierCheb = TensorSpace{<:Tuple{<:Fourier,<:Chebyshev}}
const LaurentCheb = TensorSpace{<:Tuple{<:Laurent,<:Chebyshev}}
function tensor_execute!(kind::Cint, M::Matrix{T}) where T
    pl = B200Plan{T}(kind, size(M, 1), 1, size(M, 1); n2 = size(M, 2))
    GC.@preserve M check(ccall((:afb_plan_execute_host, LIB), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}), pl.handle, M, M))
    M
end
transform!(::ChebCheb, M::Matrix{Float64}) = tensor_execute!(CHEB1_2D_FWD, M)
itransform!(::ChebCheb, M::Matrix{Float64}) = tensor_execute!(CHEB1_2D_INV, M)
transform!(::FourierCheb, M::Matrix{Float64}) = tensor_execute!(FOURIER_CHEB1_2D_FWD, M)
itransform!(::FourierCheb, M::Matrix{Float64}) = tensor_execute!(FOURIER_CHEB1_2D_INV, M)
transform!(::LaurentCheb, M::Matrix{ComplexF64}) = tensor_execute!(LAURENT_CHEB1_2D_FWD, M)
itransform!(::LaurentCheb, M::Matrix{ComplexF64}) = tensor_execute!(LAURENT_CHEB1_2D_INV, M)

# ---- Clenshaw ------------------------------------------------------------------------------------------------------
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
# clenshaw(sp, c, x::Vector) for point sets (the scalar method, src/Extras/sample.jl:47, stays in Julia: a GPU round trip
# per scalar costs ~50 us)
clenshaw(sp::Chebyshev, c::Vector{Float64}, x::Vector{Float64}) = clenshaw(c, x, ClenshawPlan(Float64, sp, length(c), length(x)))
# general three-term recurrence spaces: p_{j+1} = (A_j x + B_j) p_j - C_j p_{j-1} with the upstream recA / recB / recC tables
function clenshaw(sp::Union{Jacobi,Ultraspherical}, c::Vector{Float64}, x::Vector{Float64})
    N = length(c)
    A = Float64[recA(Float64, sp, j) for j in 0:N-1]
    B = Float64[recB(Float64, sp, j) for j in 0:N-1]
    Cc = Float64[recC(Float64, sp, j) for j in 0:N]
    y = similar(x)
    check(ccall((:afb_clenshaw_rec, LIB), Int32,
                (Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                c, N, A, B, Cc, x, y, length(x)))
    y
end

# f(x_p, y_p) for a Chebyshev x Chebyshev / Fourier x Chebyshev coefficient matrix (canonical variables)
function clenshaw2d(C::Matrix{Float64}, x::Vector{Float64}, y::Vector{Float64}; fourier::Bool = false)
    length(x) == length(y) || throw(DimensionMismatch("x and y differ in length"))
    out = similar(x)
    if fourier
        check(ccall((:afb_clenshaw_fourier_cheb2d, LIB), Int32, (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                    C, size(C, 1), size(C, 2), x, y, out, length(x)))
    else
        check(ccall((:afb_clenshaw_cheb2d, LIB), Int32, (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                    C, size(C, 1), size(C, 2), x, y, out, length(x)))
    end
    out
end

# ---- sampling (src/Extras/sample.jl) -----------------------------------------------------------------------------
# chebbisectioninv (sample.jl:55-101): the 47 fused Clenshaw bisections
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

# chebnormalizedcumsum! (sample.jl:170-175)
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

# bisectioninv for Chebyshev Funs (sample.jl:103-109) with fromcanonical fused in the kernel
function bisectioninv(cf::Fun{<:Chebyshev,Float64}, x::Vector{Float64})
    d = domain(cf)
    out = similar(x)
    c = coefficients(cf)
    check(ccall((:afb_samplecdf_cheb, LIB), Int32,
                (Ptr{Float64}, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Int64),
                c, length(c), leftendpoint(d), rightendpoint(d), x, out, length(x)))
    out
end
# the GENERIC bisectioninv(f::Fun, x::AbstractVector; numits = 47) (sample.jl:8, 23-36) for piecewise Chebyshev Funs --
# what sample(abs(Fun(sin, -5..5)), n) uses (test/PiecewiseSampleTest.jl:5-10): one kernel instead of 47 n scalar evaluations
function bisectioninv(f::Fun{<:ApproxFun.PiecewiseSpace,Float64}, x::Vector{Float64}; numits::Int = 47)
    comps = ApproxFun.components(f)
    all(g -> space(g) isa Chebyshev, comps) || return Float64[invoke(bisectioninv, Tuple{Fun,Float64}, f, xx; numits = numits) for xx in x]
    breaks = Float64[leftendpoint(domain(comps[1])); [rightendpoint(domain(g)) for g in comps]]
    offs = Int64[0; cumsum([ncoefficients(g) for g in comps])]
    coef = reduce(vcat, [coefficients(g) for g in comps])
    out = similar(x)
    check(ccall((:afb_piecewise_bisect, LIB), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                coef, offs, breaks, length(comps), x, out, length(x), numits))
    out
end

# sample(f, n) with the uniforms drawn on the device (Philox4x32-10 counter i for sample i): no rand(n), no H2D traffic
function sample_b200(f::Fun{<:Chebyshev,Float64}, n::Integer; seed::Integer=0)
    cf = ApproxFun.normalizedcumsum(f)
    d = domain(f)
    c = coefficients(cf)
    out = Vector{Float64}(undef, n)
    check(ccall((:afb_sample_cheb, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, UInt64, Ptr{Float64}, Int64),
                c, length(c), leftendpoint(d), rightendpoint(d), seed, out, n))
    out
end

# lines 208-213 of src/Extras/sample.jl in one call:  fA = evaluate.(f.A, ry'); AB = CB*fA; chebnormalizedcumsum!(AB);
# rx = chebbisectioninv(AB, u); fromcanonical.(domain(f,1), rx)   (any number of coefficients in f.B)
function sample_conditional(A::Matrix{Float64}, CB::Matrix{Float64}, dA, d1, ry::Vector{Float64}, u::Vector{Float64})
    size(A, 2) == size(CB, 2) || throw(DimensionMismatch("rank mismatch"))
    rx = similar(u)
    check(ccall((:afb_sample2d_cheb_lowrank, LIB), Int32,
                (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Float64, Float64, Float64, Float64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                A, size(A, 1), CB, size(CB, 1), size(A, 2), leftendpoint(dA), rightendpoint(dA),
                leftendpoint(d1), rightendpoint(d1), ry, u, rx, length(u)))
    rx
end

# sibling plans on the same engine: kinds 8/9 = FastTransforms Val(2) (REDFT00), 10/11 = SinSpace, 12..15 = Taylor/Hardy
chebyshevtransform2(v::VecOrMat{Float64}) = B200Plan{Float64}(Cint(8), size(v, 1), size(v, 2)) * v
ichebyshevtransform2(c::VecOrMat{Float64}) = B200Plan{Float64}(Cint(9), size(c, 1), size(c, 2)) * c

# Padua: points(Chebyshev()^2, N) and the vector transform of Fun(f, Chebyshev()^2, N)  (kinds 16 / 17)
function paduapoints_b200(N::Integer, dx, dy)
    np = Ref{Int64}(0)
    check(ccall((:afb_paduapoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Ref{Int64}),
                C_NULL, N, leftendpoint(dx), rightendpoint(dx), leftendpoint(dy), rightendpoint(dy), np))
    xy = Matrix{Float64}(undef, np[], 2)
    check(ccall((:afb_paduapoints, LIB), Int32, (Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Ref{Int64}),
                xy, N, leftendpoint(dx), rightendpoint(dx), leftendpoint(dy), rightendpoint(dy), np))
    xy
end
paduatransform_b200(v::Vector{Float64}) = B200Plan{Float64}(Cint(16), length(v), 1) * v
ipaduatransform_b200(c::Vector{Float64}) = B200Plan{Float64}(Cint(17), length(c), 1) * c

# ---- device-resident buffers (afb_buf_*): values -> point-wise op -> transform chains without PCIe round trips ----------
mutable struct B200Buf
    handle::Ptr{Cvoid}
    n::Int
    function B200Buf(n::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:afb_buf_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int64), h, n))
        b = new(h[], n)
        finalizer(q -> ccall((:afb_buf_destroy, LIB), Int32, (Ptr{Cvoid},), q.handle), b)
        b
    end
end
upload(v::Vector{Float64}) = (b = B200Buf(length(v)); check(ccall((:afb_buf_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), b.handle, v, length(v), 0)); b)
download(b::B200Buf, n::Integer = b.n) = (v = Vector{Float64}(undef, n); check(ccall((:afb_buf_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), b.handle, v, n, 0)); v)
const UNARY = Dict(:identity => 0, :sin => 1, :cos => 2, :exp => 3, :abs => 4, :sqrt => 5, :abs2 => 6, :inv => 7, :log => 8, :tanh => 9)
function unary(op::Symbol, a::B200Buf; alpha = 1.0, beta = 0.0)      # op.(alpha .* a .+ beta)
    out = B200Buf(a.n)
    check(ccall((:afb_buf_unary, LIB), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float64, Float64), UNARY[op], a.handle, out.handle, a.n, alpha, beta))
    out
end
function binary(op::Symbol, a::B200Buf, b::B200Buf)                  # a .+ b, .-, .*, ./
    out = B200Buf(a.n)
    code = Dict(:+ => 0, :- => 1, :* => 2, :/ => 3)[op]
    check(ccall((:afb_buf_binary, LIB), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), code, a.handle, b.handle, out.handle, a.n))
    out
end
function chebpoints_device(sp::Chebyshev, n::Integer)
    d = domain(sp)
    b = B200Buf(n)
    check(ccall((:afb_buf_chebpoints, LIB), Int32, (Ptr{Cvoid}, Int64, Float64, Float64), b.handle, n, leftendpoint(d), rightendpoint(d)))
    b
end
function transform_device(p::B200Plan, v::B200Buf)
    out = B200Buf(v.n)
    check(ccall((:afb_plan_execute_buf, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.handle, v.handle, out.handle))
    out
end
# e.g.  x = chebpoints_device(S, n); c = download(transform_device(B200Plan{Float64}(CHEB1_FWD, n, 1), unary(:sin, unary(:abs2, x))))
# is Fun(x -> sin(x^2), S, n).coefficients with no host array until the download.

end # module
