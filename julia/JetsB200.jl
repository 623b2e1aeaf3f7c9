# JetsB200.jl -- the reference-side binding for libjets_b200.so.
#
# STATUS: EXPERIMENTAL / UNEXECUTED.  Julia is not installed in the build image, so this file has never
# been parsed or run; it is the stub a Jets.jl maintainer would add (see INTEGRATION.md) and must be brought
# up under Julia CI before use.  It is kept thin so that it can be checked by inspection against
# include/jets_b200.h (tests/test_julia_shim_static.py checks every ccall's name, arity and argument classes).
# Every `ccall` below names one entry point of that header; nothing else crosses the boundary.
#
# What it provides, in Jets' own vocabulary (src/Jets.jl line numbers of v1.4.1):
#   * B200Array{T} <: AbstractVector{T}   -- device storage for JetSpace / JetBSpace vectors
#     (flat buffer + block offset table; `getblock` is a view).           :105-108, :809-924
#   * B200 leaf constructors returning ordinary JopLn / JopNl whose closures ccall the
#     library: JopDiagonalB200, JopPointwiseB200, JopStencilB200, JopDenseB200.
#   * `LinearAlgebra.mul!` methods for JopLn/JopNl/JopAdjoint whose tree is all-B200: the WHOLE
#     tree (block / sum / composite / adjoint) is handed to ONE `jets_apply`.  :390-392
#   * device methods for dot / norm / fill! / extrema / broadcast-axpy.   :834-911
#   * DistJop: the rank-local rows of a block-row partitioned JopBlock (one process per GPU) -- `mul!` is ONE
#     `jets_dist_apply`, the halo exchange happens inside the kernel; `normal_host!` is the host-buffer pipeline.
#   * device scalars, fused apply-axpby and CUDA-graph capture for CG/LSQR loops without host round trips.
module JetsB200

using Jets, LinearAlgebra

const LIB = get(ENV, "JETS_B200_LIB", "libjets_b200")

struct JetsB200Error <: Exception
    code::Cint
    msg::String
end
function check(code::Cint)
    code == 0 && return nothing
    throw(JetsB200Error(code, unsafe_string(ccall((:jets_last_error, LIB), Cstring, ()))))
end

init(device::Integer = 0) = check(ccall((:jets_init, LIB), Cint, (Cint,), device))
dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)
dtype_code(::Type{ComplexF32}) = Cint(2)      # JETS_C64: interleaved (re, im), Julia's own layout
dtype_code(::Type{ComplexF64}) = Cint(3)      # JETS_C128

# ------------------------------------------------------------------ device storage --------
mutable struct B200Array{T} <: AbstractVector{T}
    h::Ptr{Cvoid}                 # jets_buf
    blocklengths::Vector{Int64}
    function B200Array{T}(h::Ptr{Cvoid}, bl::Vector{Int64}) where {T}
        x = new{T}(h, bl)
        finalizer(x -> ccall((:jets_buf_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), x)
        x
    end
end
Base.size(x::B200Array) = (sum(x.blocklengths),)

# zeros(R) / Array(R) for JetSpace and JetBSpace                      src/Jets.jl:105-108, :922-924
blocklengths(R::Jets.JetSpace) = Int64[length(R)]
blocklengths(R::Jets.JetBSpace) = Int64[length(Jets.space(R, i)) for i in 1:Jets.nblocks(R)]
function b200zeros(R::Jets.JetAbstractSpace{T}) where {T}
    bl = blocklengths(R)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_buf_create, LIB), Cint, (Cint, Int32, Ptr{Int64}, Ptr{Ptr{Cvoid}}),
                dtype_code(T), length(bl), bl, h))
    B200Array{T}(h[], bl)
end
# similar / copy / zeros on the device: Jets' own `jacobian` runs copy(mₒ) (:374) and `A*m` runs
# zeros(range(A)) (:399); neither may fall back to AbstractArray's element-by-element generics.
function Base.similar(x::B200Array{T}) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_buf_create, LIB), Cint, (Cint, Int32, Ptr{Int64}, Ptr{Ptr{Cvoid}}),
                dtype_code(T), length(x.blocklengths), x.blocklengths, h))
    B200Array{T}(h[], copy(x.blocklengths))
end
function Base.copyto!(dst::B200Array{T}, src::B200Array{T}) where {T}
    check(ccall((:jets_buf_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dst.h, src.h)); dst
end
Base.copy(x::B200Array) = copyto!(similar(x), x)
function b200rand(R::Jets.JetAbstractSpace, seed::Integer = rand(UInt64) >> 1)
    x = b200zeros(R)
    check(ccall((:jets_buf_rand, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Cint), x.h, seed, 0, 0)); x
end
# host <-> device: setblock!/getblock!/convert(Array,x)               src/Jets.jl:862-868, :915-916
function Base.copyto!(x::B200Array{T}, a::Array{T}) where {T}
    check(ccall((:jets_buf_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{T}, Int64), x.h, -1, a, length(a)))
    x
end
function Base.Array(x::B200Array{T}) where {T}
    a = Vector{T}(undef, length(x))
    check(ccall((:jets_buf_download, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{T}, Int64), x.h, -1, a, length(a)))
    a
end
function Jets.getblock(x::B200Array{T}, i::Integer) where {T}              # a view, :914
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_buf_view, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}), x.h, i - 1, 1, h))
    B200Array{T}(h[], [x.blocklengths[i]])
end
function Jets.setblock!(x::B200Array{T}, i::Integer, a::Array{T}) where {T} # :916
    check(ccall((:jets_buf_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{T}, Int64), x.h, i - 1, a, length(a)))
end

# reductions and updates                                               src/Jets.jl:834-911
function LinearAlgebra.dot(x::B200Array{T}, y::B200Array{T}) where {T}
    r = Ref{Cdouble}(0)
    check(ccall((:jets_dot, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}), x.h, y.h, r))
    T(r[])
end
# complex eltypes: dot conjugates its first argument (:853) and returns a complex number
function LinearAlgebra.dot(x::B200Array{T}, y::B200Array{T}) where {T<:Complex}
    r = zeros(Cdouble, 2)
    check(ccall((:jets_dot_c, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}), x.h, y.h, r))
    T(r[1], r[2])
end
function LinearAlgebra.norm(x::B200Array{T}, p::Real = 2) where {T}
    r = Ref{Cdouble}(0)
    check(ccall((:jets_norm, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), x.h, p, r))
    real(T)(r[])
end
# norm of a SymmetricArray whose parent lives on the device (src/Jets.jl:443-462): `w` holds, per stored
# element, 1 + the number of mirrored positions that map onto it (computed once per JetSSpace on the host)
function symmetric_norm(parent::B200Array{T}, w::B200Array{Float64}, p::Real = 2) where {T<:Complex}
    r = Ref{Cdouble}(0)
    check(ccall((:jets_norm_weighted, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), parent.h, w.h, p, r))
    real(T)(r[])
end
# abs.(x) of a complex vector: a real vector on the same block structure (test/runtests.jl:545-547)
function absb200!(out::B200Array{R}, x::B200Array{Complex{R}}) where {R}
    check(ccall((:jets_abs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), out.h, x.h)); out
end
Base.fill!(x::B200Array, a) = (check(ccall((:jets_buf_fill_c, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), x.h, real(a), imag(a))); x)
# scalar getindex / setindex! (src/Jets.jl:819-832): one element over the host link, 1-based
function Base.getindex(x::B200Array{T}, i::Integer) where {T}
    v = Vector{T}(undef, 1)
    check(ccall((:jets_buf_read, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{T}, Int64), x.h, i - 1, v, 1)); v[1]
end
function Base.setindex!(x::B200Array{T}, a, i::Integer) where {T}
    check(ccall((:jets_buf_write, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{T}, Int64), x.h, i - 1, T[a], 1)); a
end
function Base.extrema(x::B200Array{T}) where {T}
    mn = Ref{Cdouble}(0); mx = Ref{Cdouble}(0)
    check(ccall((:jets_extrema, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), x.h, mn, mx))
    (T(mn[]), T(mx[]))
end
# out .= a.*x .+ b.*y  (the axpy-class broadcasts of CG/LSQR; :905-911)
function axpby!(out::B200Array, a::Real, x::B200Array, b::Real, y::B200Array)
    c = Cdouble[a, b]; hs = Ptr{Cvoid}[x.h, y.h]
    check(ccall((:jets_lincomb, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}), out.h, 2, c, hs))
    out
end

# ------------------------------------------------------------------ operators --------------
# The jets_op handle lives in the Jet's state NamedTuple (`s.b200`), so Jets' own combinators
# (∘, +, -, @blockop, adjoint, jacobian) keep working unchanged; `handle(A)` maps a Jets tree onto
# a library tree once and caches it.
mutable struct OpHandle
    h::Ptr{Cvoid}
    function OpHandle(h)
        x = new(h)
        finalizer(x -> ccall((:jets_op_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), x)
        x
    end
end
# Jets' own `jacobian(F, mₒ)` runs `copy(jet, false)` = `deepcopy(jet.s)` (src/Jets.jl:230, :374): a device
# handle must not be copied bit for bit (two finalizers, one pointer).  An operator handle is cloned (new
# nodes, shared immutable state: a later point! on the copy leaves the original alone); a device vector in
# the state (a diagonal) is shared by reference count.
function Base.deepcopy_internal(x::OpHandle, d::IdDict)
    haskey(d, x) && return d[x]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_clone, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), x.h, h))
    d[x] = OpHandle(h[])
end
function Base.deepcopy_internal(x::B200Array{T}, d::IdDict) where {T}
    haskey(d, x) && return d[x]
    check(ccall((:jets_buf_retain, LIB), Cint, (Ptr{Cvoid},), x.h))
    d[x] = B200Array{T}(x.h, copy(x.blocklengths))
end
# The closures exist so that a B200 leaf is a legal Jet; they are reached when a tree mixes B200 leaves with
# other operators that accept B200Arrays (mul! below falls back to Jets' own dispatch for such trees, and each
# B200 leaf is then one jets_apply on its own).
function leaf_apply!(out::B200Array, h::OpHandle, mode::Integer, in::B200Array)
    check(ccall((:jets_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint), h.h, mode, out.h, in.h, 0))
    out
end

# JopLn(df! = d .= w.*m)  -- fixture JopFoo, test/runtests.jl:3-8
function JopDiagonalB200(w::B200Array{T}) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_diag, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), w.h, h))
    oh = OpHandle(h[])
    sp = JetSpace(T, length(w))
    JopLn(dom = sp, rng = sp, s = (b200 = oh, w = w),
          df! = (d, m; b200, kw...) -> leaf_apply!(d, b200, 1, m),
          df′! = (m, d; b200, kw...) -> leaf_apply!(m, b200, 2, d))
end
# d = m[indices], adjoint m .= 0; m[indices] = d -- the JetPack-style restriction (indices 1-based, unique)
function JopRestrictionB200(::Type{T}, n::Integer, indices::AbstractVector{<:Integer}) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    idx0 = Int64.(indices) .- 1
    check(ccall((:jets_op_restrict, LIB), Cint, (Cint, Int64, Int64, Ptr{Int64}, Ptr{Ptr{Cvoid}}), dtype_code(T), n, length(idx0), idx0, h))
    oh = OpHandle(h[])
    JopLn(dom = JetSpace(T, n), rng = JetSpace(T, length(idx0)), s = (b200 = oh, indices = indices),
          df! = (d, m; b200, kw...) -> leaf_apply!(d, b200, 1, m),
          df′! = (m, d; b200, kw...) -> leaf_apply!(m, b200, 2, d))
end
# JopNl(f! = phi(m), df! = phi'(mo).*dm) -- fixture JopBar, test/runtests.jl:20-25
function JopPointwiseB200(::Type{T}, n::Integer, fn::Integer = 0, p::Real = 0) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_pointwise, LIB), Cint, (Cint, Int64, Cint, Cdouble, Ptr{Ptr{Cvoid}}), dtype_code(T), n, fn, p, h))
    oh = OpHandle(h[])
    # jets_apply refuses mode DFT on a nonlinear handle (src/Jets.jl:392): df!/df′! go through the linear view of
    # the SAME jet (jets_op_as_linear: it sees the point that upstate! sets on `oh`)
    lin = unary(:jets_op_as_linear, oh)
    sp = JetSpace(T, n)
    JopNl(dom = sp, rng = sp, s = (b200 = oh, b200lin = lin),
          f! = (d, m; b200, kw...) -> leaf_apply!(d, b200, 0, m),
          df! = (d, m; mₒ, b200lin, kw...) -> leaf_apply!(d, b200lin, 1, m),
          df′! = (m, d; mₒ, b200lin, kw...) -> leaf_apply!(m, b200lin, 2, d),
          upstate! = (mₒ, s) -> check(ccall((:jets_op_set_point, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.b200.h, mₒ.h)))
end
function JopStencilB200(::Type{T}, n::Integer, kind::Integer = 0) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_stencil, LIB), Cint, (Cint, Int64, Cint, Ptr{Ptr{Cvoid}}), dtype_code(T), n, kind, h))
    oh = OpHandle(h[])
    sp = JetSpace(T, n)
    JopLn(dom = sp, rng = sp, s = (b200 = oh,),
          df! = (d, m; b200, kw...) -> leaf_apply!(d, b200, 1, m),
          df′! = (m, d; b200, kw...) -> leaf_apply!(m, b200, 2, d))
end
# Matrix as operator (src/Jets.jl:325-326, :573-576); A is a B200Array holding rows*cols column-major
function JopDenseB200(A::B200Array{T}, rows::Integer, cols::Integer; nrhs::Integer = 1) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_dense, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}), A.h, rows, cols, nrhs, h))
    oh = OpHandle(h[])
    dom = nrhs == 1 ? JetSpace(T, cols) : JetSpace(T, cols, nrhs)
    rng = nrhs == 1 ? JetSpace(T, rows) : JetSpace(T, rows, nrhs)
    JopLn(dom = dom, rng = rng, s = (b200 = oh, A = A),
          df! = (d, m; b200, kw...) -> leaf_apply!(d, b200, 1, m),
          df′! = (m, d; b200, kw...) -> leaf_apply!(m, b200, 2, d))
end

# ---- tree -> handle (built once per Jet, cached in a WeakKeyDict) ---------------------------
const TREE = WeakKeyDict{Any,OpHandle}()
isb200(j::Jets.Jet) = haskey(state(j), :b200) ||
    (haskey(state(j), :ops) && all(op -> isb200(jet(op)), state(j).ops))

function handle(A::Jets.Jop)
    A isa Jets.JopAdjoint && return unary(:jets_op_adjoint, handle(A.op))
    j = jet(A)
    base = get!(TREE, j) do
        s = state(j)
        if haskey(s, :b200)
            s.b200
        elseif j.f! === Jets.JetComposite_f!                               # :522-576
            nary(:jets_op_compose, [handle(op) for op in s.ops])
        elseif j.f! === Jets.JetSum_f!                                     # :628-708
            hs = [handle(op) for op in s.ops]; sg = Int32[sgn == (+) ? 1 : -1 for sgn in s.sgns]
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:jets_op_sum, LIB), Cint, (Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
                        length(hs), [x.h for x in hs], sg, h))
            OpHandle(h[])
        elseif j.f! === Jets.JetBlock_f!                                   # :926-1057, column-major like Julia
            hs = [handle(op) for op in s.ops]                              # Matrix{Jop} iterates column-major
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:jets_op_block, LIB), Cint, (Int32, Int32, Ptr{Ptr{Cvoid}}, Cint, Ptr{Ptr{Cvoid}}),
                        size(s.ops, 1), size(s.ops, 2), [x.h for x in hs], s.dom isa Jets.JetBSpace && size(s.ops, 2) == 1, h))
            OpHandle(h[])
        else
            error("not a B200 operator tree")
        end
    end
    # a JopLn over a nonlinear jet applies df! (the linear view), :209-224
    (A isa Jets.JopLn && !islinear(base)) ? unary(:jets_op_as_linear, base) : base
end
islinear(h::OpHandle) = ccall((:jets_op_is_linear, LIB), Cint, (Ptr{Cvoid},), h.h) == 1
function unary(f::Symbol, a::OpHandle)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    code = f === :jets_op_adjoint ?
        ccall((:jets_op_adjoint, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), a.h, h) :
        ccall((:jets_op_as_linear, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), a.h, h)
    check(code)
    OpHandle(h[])
end
function nary(::Symbol, hs::Vector{OpHandle})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jets_op_compose, LIB), Cint, (Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}), length(hs), [x.h for x in hs], h))
    OpHandle(h[])
end

# ---- mul! overrides: the whole tree in ONE call (src/Jets.jl:390-392) -------------------------
const MODE_F, MODE_DF, MODE_DFT = Cint(0), Cint(1), Cint(2)
apply!(d::B200Array, A::Jets.Jop, m::B200Array, mode::Cint) =
    (check(ccall((:jets_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint), handle(A).h, mode, d.h, m.h, 0)); d)
# all-B200 trees go to the library whole; anything else takes Jets' own closure dispatch (:390-392)
LinearAlgebra.mul!(d::B200Array, A::Jets.JopNl, m::B200Array) =
    isb200(jet(A)) ? apply!(d, A, m, MODE_F) : invoke(mul!, Tuple{AbstractArray,Jets.JopNl,AbstractArray}, d, A, m)
LinearAlgebra.mul!(d::B200Array, A::Jets.JopLn, m::B200Array) =
    isb200(jet(A)) ? apply!(d, A, m, MODE_DF) : invoke(mul!, Tuple{AbstractArray,Jets.JopLn,AbstractArray}, d, A, m)
LinearAlgebra.mul!(m::B200Array, A::Jets.JopAdjoint, d::B200Array) =
    isb200(jet(A)) ? apply!(m, A.op, d, MODE_DFT) : invoke(mul!, Tuple{AbstractArray,Jets.JopAdjoint,AbstractArray}, m, A, d)
Base.:*(A::Jets.Jop, m::B200Array) = mul!(b200zeros(range(A)), A, m)       # :399

# point! of a composition evaluates the chain with `mul!(zeros(range(ops[i])), ops[i], _m)` (:578-589); zeros(R)
# is a host Array, so the device version allocates on the device instead.  (Sums and block operators recurse
# through point!/getblock only, :710-715 and :1059-1066, and need nothing.)
function Jets.point!(j::Jets.Jet{D,R,typeof(Jets.JetComposite_f!)}, mₒ::B200Array) where {D<:Jets.JetAbstractSpace,R<:Jets.JetAbstractSpace}
    j.mₒ = mₒ
    ops = state(j).ops
    _m = copy(mₒ)
    for i = length(ops):-1:1
        Jets.point!(jet(ops[i]), _m)
        if i > 1
            _m = mul!(b200zeros(range(ops[i])), ops[i], _m)
        end
    end
    j
end

# ------------------------------------------------------------------ device scalars, fused updates, graphs ----
# A CG/LSQR iteration without a host round trip (docs/src/index.md:235-246): reductions leave their result on
# the device, updates read their coefficients from there, and the whole iteration is captured in a CUDA graph.
mutable struct B200Scalar
    h::Ptr{Cvoid}
    function B200Scalar(v::Real = 0.0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:jets_scalar_create, LIB), Cint, (Ptr{Ptr{Cvoid}},), h))
        s = new(h[])
        finalizer(s -> ccall((:jets_scalar_destroy, LIB), Cint, (Ptr{Cvoid},), s.h), s)
        check(ccall((:jets_scalar_set, LIB), Cint, (Ptr{Cvoid}, Cdouble), s.h, v))
        s
    end
end
function Base.getindex(s::B200Scalar)
    v = Ref{Cdouble}(0)
    check(ccall((:jets_scalar_get, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, v)); v[]
end
dot!(out::B200Scalar, x::B200Array, y::B200Array) =
    (check(ccall((:jets_dot_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), x.h, y.h, out.h)); out)
norm!(out::B200Scalar, x::B200Array, p::Real = 2) =
    (check(ccall((:jets_norm_dev, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), x.h, p, out.h)); out)
# out[i] = a[i] (op[i]) b[i] in ONE launch; op in '+','-','*','/', 'n' (negate), 's' (sqrt), 'h' (hypot)
function scalar_prog!(outs::Vector{B200Scalar}, ops::String, as::Vector{B200Scalar}, bs::Vector{<:Union{B200Scalar,Nothing}})
    nul(x) = x === nothing ? C_NULL : x.h
    check(ccall((:jets_scalar_prog, LIB), Cint, (Int32, Ptr{Ptr{Cvoid}}, Cstring, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                length(outs), [o.h for o in outs], ops, [a.h for a in as], [nul(b) for b in bs]))
end
# out .= (sa or ca) .* x .+ (sb or cb) .* y with device-resident coefficients; flags: 1 negate, 2 reciprocal
function axpby!(out::B200Array, sa::Union{B200Scalar,Nothing}, ca::Real, aflags::Integer, x::B200Array,
                sb::Union{B200Scalar,Nothing}, cb::Real, bflags::Integer, y::B200Array)
    nul(s) = s === nothing ? C_NULL : s.h
    check(ccall((:jets_axpby_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}),
                out.h, nul(sa), ca, aflags, x.h, nul(sb), cb, bflags, y.h))
    out
end
# out .= cA .* (A in) .+ cO .* out in ONE pass (the Golub-Kahan updates u = A v - alpha u, v = A'u - beta v)
function mul_axpby!(out::B200Array, A::Jets.Jop, in::B200Array, sa::Union{B200Scalar,Nothing}, ca::Real, aflags::Integer,
                    so::Union{B200Scalar,Nothing}, co::Real, oflags::Integer)
    nul(s) = s === nothing ? C_NULL : s.h
    mode = A isa Jets.JopAdjoint ? MODE_DFT : MODE_DF
    B = A isa Jets.JopAdjoint ? A.op : A
    check(ccall((:jets_apply_axpby, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}, Cdouble, Cint),
                handle(B).h, mode, out.h, in.h, nul(sa), ca, aflags, nul(so), co, oflags))
    out
end
# the same, and nrm = norm(out) of the vector just written (LSQR's beta = ||u||, alpha = ||v||), one call
function mul_axpby_norm!(out::B200Array, A::Jets.Jop, in::B200Array, sa::Union{B200Scalar,Nothing}, ca::Real, aflags::Integer,
                         so::Union{B200Scalar,Nothing}, co::Real, oflags::Integer, nrm::B200Scalar)
    nul(s) = s === nothing ? C_NULL : s.h
    mode = A isa Jets.JopAdjoint ? MODE_DFT : MODE_DF
    B = A isa Jets.JopAdjoint ? A.op : A
    check(ccall((:jets_apply_axpby_norm, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}),
                handle(B).h, mode, out.h, in.h, nul(sa), ca, aflags, nul(so), co, oflags, nrm.h))
    out
end
# two axpby updates in ONE pass (CG: x += a p, r -= a q; LSQR: x += t1 w, w = v/alpha - t2 w); outputs may alias inputs
function axpby_pair!(out1::B200Array, s1a, c1a::Real, f1a::Integer, x1::B200Array, s1b, c1b::Real, f1b::Integer, y1::B200Array,
                     out2::B200Array, s2a, c2a::Real, f2a::Integer, x2::B200Array, s2b, c2b::Real, f2b::Integer, y2::B200Array)
    nul(s) = s === nothing ? C_NULL : s.h
    check(ccall((:jets_axpby_pair_dev, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid},
                 Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}),
                out1.h, nul(s1a), c1a, f1a, x1.h, nul(s1b), c1b, f1b, y1.h, out2.h, nul(s2a), c2a, f2a, x2.h, nul(s2b), c2b, f2b, y2.h))
end
# graph = capture() do ... library calls ... end; launch(graph) replays them
mutable struct B200Graph
    h::Ptr{Cvoid}
end
function capture(f)
    check(ccall((:jets_graph_begin, LIB), Cint, ()))
    g = Ref{Ptr{Cvoid}}(C_NULL)
    try
        f()
    finally
        check(ccall((:jets_graph_end, LIB), Cint, (Ptr{Ptr{Cvoid}},), g))
    end
    x = B200Graph(g[])
    finalizer(x -> ccall((:jets_graph_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), x)
    x
end
launch(g::B200Graph) = check(ccall((:jets_graph_launch, LIB), Cint, (Ptr{Cvoid},), g.h))
sync() = check(ccall((:jets_sync, LIB), Cint, ()))

# ------------------------------------------------------------------ distributed operators --------------------
# One process per GPU.  The communicator is bootstrapped either through NCCL (unique id broadcast by the host
# program, e.g. MPI.Bcast) or through a host all-gather callback (MPI.Allgather behind @cfunction): the banded
# path needs only the latter -- its payload moves through peer memory inside the apply kernel.
dist_unique_id() = (id = zeros(UInt8, 128); check(ccall((:jets_dist_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
dist_init(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ccall((:jets_dist_init, LIB), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id))
# allgather = @cfunction(my_allgather, Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64))
dist_init_host(rank::Integer, nranks::Integer, allgather::Ptr{Cvoid}, user::Ptr{Cvoid} = C_NULL) =
    check(ccall((:jets_dist_init_host, LIB), Cint, (Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}), rank, nranks, allgather, user))
dist_shutdown() = check(ccall((:jets_dist_shutdown, LIB), Cint, ()))
function dist_sum(v::Real)      # rank-ordered (bit-stable) sum of a host scalar: distributed dot / norm
    r = Ref{Cdouble}(v)
    check(ccall((:jets_dist_sum_scalar, LIB), Cint, (Ptr{Cdouble},), r)); r[]
end

# The rank-local rows of a block-row partitioned JopBlock.  `Aloc` is an ordinary all-B200 JopBlock: nloc x
# (nloc + 2*halo) over [halo blocks of rank-1 | own blocks | halo blocks of rank+1] (banded), or nloc x ncol
# over the whole domain (dense = true).  Vectors are the rank's own shards.
mutable struct DistJop
    h::Ptr{Cvoid}
    Aloc::Jets.Jop
    function DistJop(Aloc::Jets.Jop; halo::Integer = 1, dense::Bool = false)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if dense
            check(ccall((:jets_dist_op_create_dense, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), handle(Aloc).h, h))
        else
            check(ccall((:jets_dist_op_create, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}), handle(Aloc).h, halo, h))
        end
        x = new(h[], Aloc)
        finalizer(x -> ccall((:jets_dist_op_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), x)
        x
    end
end
# mul!(d, A, m) / mul!(m, A', d) on the shards: ONE call, one kernel launch per rank (banded)
LinearAlgebra.mul!(d::B200Array, A::DistJop, m::B200Array) =
    (check(ccall((:jets_dist_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), A.h, MODE_DF, d.h, m.h)); d)
mul_adjoint!(m::B200Array, A::DistJop, d::B200Array) =
    (check(ccall((:jets_dist_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), A.h, MODE_DFT, m.h, d.h)); m)
# host_out = A'(A host_in) for this rank's shards, chunk-pipelined inside the library (pinned host arrays)
function normal_host!(host_out::Vector{T}, A::DistJop, host_in::Vector{T}; nchunks::Integer = 0) where {T}
    check(ccall((:jets_dist_apply_normal_host, LIB), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Int32), A.h, host_out, host_in, nchunks))
    host_out
end
join!(A::DistJop) = check(ccall((:jets_dist_op_join, LIB), Cint, (Ptr{Cvoid},), A.h))
# collective: lets the neighbours map the domain shard x, so that forward applies on it read their halo blocks in place
register!(A::DistJop, x::B200Array) = (check(ccall((:jets_dist_op_register, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), A.h, x.h)); x)

end # module
