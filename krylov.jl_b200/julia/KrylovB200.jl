# KrylovB200.jl -- the Julia face of the boundary: a storage type + operator type for which
# Krylov.jl's documented plugin points (docs/src/custom_workspaces.md:107,151-300) resolve to
# `ccall`s into libkrylov_b200.so.
#
# STATUS: Julia is not installed in the build image, so this file has never been executed here.
# It is the reference-side binding a maintainer would add (see INTEGRATION.md); everything it
# calls is exercised through the same C ABI by tests/ via ctypes.
#
# Two levels, as in the reference:
#   (1) primitive level  -- Krylov.kdot / knorm / kaxpy! / ... / kmul! overloads for B200Vector /
#       B200CSR.  Every solver of Krylov.jl then runs on the GPU unmodified (scalars come back to
#       the host by value, exactly like CuVector storage, src/krylov_utils.jl:309-349).
#   (2) solver level     -- cg!/gmres!/bicgstab!/minres! methods for workspaces whose storage is
#       B200Vector: one C call per solve, fused kernels, device-resident scalars.
module KrylovB200

using Krylov, LinearAlgebra, SparseArrays
import Krylov: kdot, kdotr, knorm, kscal!, kdiv!, kaxpy!, kaxpby!, kcopy!, kscalcopy!, kdivcopy!, kfill!, kmul!

const lib = get(ENV, "KRYLOV_B200_LIB", "libkrylov_b200.so")
const BlasT = Union{Float32, Float64}
dtype_id(::Type{Float32}) = Cint(0)
dtype_id(::Type{Float64}) = Cint(1)

# ---- execution context (one stream + reduction scratch), lazily created -------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
  if CTX[] == C_NULL
    CTX[] = ccall((:kb200_ctx_create, lib), Ptr{Cvoid}, (Cint,), -1)
    CTX[] == C_NULL && error(unsafe_string(ccall((:krylov_b200_last_error, lib), Cstring, ())))
  end
  CTX[]
end
check(rc) = rc == 0 ? nothing : error(unsafe_string(ccall((:krylov_b200_last_error, lib), Cstring, ())))

# ---- storage type: S <: AbstractVector{T} with length/similar/S(undef,n) (inplace.md:29-38) ------
mutable struct B200Vector{T<:BlasT} <: AbstractVector{T}
  ptr::Ptr{Cvoid}
  n::Int
  function B200Vector{T}(::UndefInitializer, n::Integer) where T
    p = n == 0 ? C_NULL : ccall((:kb200_alloc, lib), Ptr{Cvoid}, (Clonglong,), n * sizeof(T))
    v = new{T}(p, n)
    finalizer(x -> (x.ptr != C_NULL && ccall((:kb200_free, lib), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), v)
    v
  end
end
function B200Vector(h::Vector{T}) where T<:BlasT
  v = B200Vector{T}(undef, length(h))
  check(ccall((:kb200_h2d, lib), Cint, (Ptr{Cvoid}, Ptr{T}, Clonglong), v.ptr, h, sizeof(h)))
  v
end
Base.size(v::B200Vector) = (v.n,)
Base.length(v::B200Vector) = v.n
Base.similar(v::B200Vector{T}) where T = B200Vector{T}(undef, v.n)
Base.similar(v::B200Vector, ::Type{T}, dims::Dims{1}) where T = B200Vector{T}(undef, dims[1])
Base.getindex(v::B200Vector, i::Int) = error("scalar indexing of a B200Vector is disabled (cf. CUDA.allowscalar(false))")
function Base.Vector(v::B200Vector{T}) where T
  h = Vector{T}(undef, v.n)
  check(ccall((:kb200_d2h, lib), Cint, (Ptr{T}, Ptr{Cvoid}, Clonglong), h, v.ptr, sizeof(h)))
  h
end

# ---- primitives: same signatures as src/krylov_utils.jl:309-347 ----------------------------------
function kdot(n::Integer, x::B200Vector{T}, y::B200Vector{T}) where T
  r = Ref{Cdouble}()
  check(ccall((:kb200_dot, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}), ctx(), dtype_id(T), n, x.ptr, y.ptr, r))
  T(r[])
end
kdotr(n::Integer, x::B200Vector{T}, y::B200Vector{T}) where T = kdot(n, x, y)
function knorm(n::Integer, x::B200Vector{T}) where T
  r = Ref{Cdouble}()
  check(ccall((:kb200_nrm2, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ref{Cdouble}), ctx(), dtype_id(T), n, x.ptr, r))
  T(r[])
end
kaxpy!(n::Integer, s::T, x::B200Vector{T}, y::B200Vector{T}) where T =
  (check(ccall((:kb200_axpy, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}), ctx(), dtype_id(T), n, s, x.ptr, y.ptr)); y)
kaxpby!(n::Integer, s::T, x::B200Vector{T}, t::T, y::B200Vector{T}) where T =
  (check(ccall((:kb200_axpby, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), ctx(), dtype_id(T), n, s, x.ptr, t, y.ptr)); y)
kscal!(n::Integer, s::T, x::B200Vector{T}) where T =
  (check(ccall((:kb200_scal, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cvoid}), ctx(), dtype_id(T), n, s, x.ptr)); x)
kdiv!(n::Integer, x::B200Vector{T}, s::T) where T = kscal!(n, one(T) / s, x)          # krylov_utils.jl:325
kcopy!(n::Integer, y::B200Vector{T}, x::B200Vector{T}) where T =
  (check(ccall((:kb200_copy, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}), ctx(), dtype_id(T), n, y.ptr, x.ptr)); y)
kscalcopy!(n::Integer, y::B200Vector{T}, s::T, x::B200Vector{T}) where T =
  (check(ccall((:kb200_scalcopy, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), ctx(), dtype_id(T), n, y.ptr, s, x.ptr)); y)
kdivcopy!(n::Integer, y::B200Vector{T}, x::B200Vector{T}, s::T) where T =
  (check(ccall((:kb200_divcopy, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), ctx(), dtype_id(T), n, y.ptr, x.ptr, s)); y)
kfill!(x::B200Vector{T}, val::T) where T =
  (check(ccall((:kb200_fill, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cdouble), ctx(), dtype_id(T), x.n, x.ptr, val)); x)

# ---- operator: CSR resident in HBM ------------------------------------------------------------------
mutable struct B200CSR{T<:BlasT}
  handle::Ptr{Cvoid}
  m::Int
  n::Int
end
# SparseMatrixCSC{T,Int64} is CSC, 1-based, Int64.  CSR(A) == CSC(A'): for the (symmetric) CG/MINRES
# operators the arrays can be passed as they are; for a general A pass the CSC arrays of copy(A').
function B200CSR(A::SparseMatrixCSC{T,Int64}; symmetric::Bool = issymmetric(A)) where T<:BlasT
  At = symmetric ? A : SparseMatrixCSC(transpose(A))
  h = ccall((:kb200_csr_create, lib), Ptr{Cvoid},
            (Ptr{Cvoid}, Cint, Cint, Clonglong, Ptr{Int64}, Ptr{Int64}, Ptr{T}, Cint, Cint, Cint),
            ctx(), dtype_id(T), size(A, 1), nnz(A), At.colptr, At.rowval, At.nzval, 1, 8, 0)
  h == C_NULL && error(unsafe_string(ccall((:krylov_b200_last_error, lib), Cstring, ())))
  op = B200CSR{T}(h, size(A)...)
  finalizer(o -> ccall((:kb200_csr_destroy, lib), Cvoid, (Ptr{Cvoid},), o.handle), op)
end
Base.size(A::B200CSR) = (A.m, A.n)
Base.eltype(::B200CSR{T}) where T = T
kmul!(y::B200Vector{T}, A::B200CSR{T}, x::B200Vector{T}) where T =       # custom_workspaces.md:114-115
  (check(ccall((:kb200_spmv_csr, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint), ctx(), A.handle, x.ptr, y.ptr, 0)); y)
LinearAlgebra.mul!(y::B200Vector, A::B200CSR, x::B200Vector) = kmul!(y, A, x)

# ---- solver level: one C call per solve (fused kernels) ------------------------------------------------
# The workspace keeps Krylov.jl's type (CgWorkspace{T,T,B200Vector{T}}) for its stats and public fields;
# the device vectors of the fused solve live in a libkrylov_b200 workspace (KRYLOV_CUDA: device pointers).
const SOLVER_ID = Dict(:cg => 0, :cr => 1, :minres => 3, :diom => 5, :dqgmres => 6, :fom => 7, :gmres => 8, :fgmres => 9,
                       :bicgstab => 10, :cgs => 11, :cg_lanczos => 100)
struct COpts   # KrylovOptions, interfaces/src/c_enums.jl:40-62
  atol::Cdouble; rtol::Cdouble; itmax::Cint; verbose::Cint; lambda::Cdouble; tau::Cdouble; nu::Cdouble
  timemax::Cdouble; radius::Cdouble; restart::Cint; reorthogonalization::Cint; linesearch::Cint
end
function fused_solve!(method::Symbol, ws, A::B200CSR{T}, b::B200Vector{T}; atol::T = √eps(T), rtol::T = √eps(T),
                      itmax::Int = 0, timemax::Float64 = Inf, verbose::Int = 0, radius::T = zero(T),
                      linesearch::Bool = false, λ::T = zero(T), restart::Bool = false,
                      reorthogonalization::Bool = false, memory::Int = 0, window::Int = 0) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  wo = (Cint(memory), Cint(window))
  rc = ccall((:krylov_workspace_create, lib), Cint, (Cint, Cint, Cint, Cint, Cint, Ref{NTuple{2,Cint}}, Ref{Ptr{Cvoid}}),
             SOLVER_ID[method], A.m, A.n, dtype_id(T), 1, wo, h)
  rc == 0 || error("krylov_workspace_create -> $rc")
  try
    # attach the operator already resident in HBM; solve; copy x (device to device) into ws.x
    check(ccall((:krylov_b200_attach_csr, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h[], A.handle))
    o = COpts(atol, rtol, itmax, verbose, λ, NaN, NaN, isinf(timemax) ? NaN : timemax, radius, restart, reorthogonalization, linesearch)
    check(ccall((:krylov_solve, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{COpts}),
               h[], C_NULL, C_NULL, C_NULL, C_NULL, b.ptr, C_NULL, C_NULL, o))
    check(ccall((:krylov_get_x, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), h[], ws.x.ptr, A.n))
    ws.stats.niter = ccall((:krylov_niter, lib), Cint, (Ptr{Cvoid},), h[])
    ws.stats.solved = ccall((:krylov_is_solved, lib), Cint, (Ptr{Cvoid},), h[]) == 1
    ws.stats.timer = ccall((:krylov_elapsed_time, lib), Cdouble, (Ptr{Cvoid},), h[])
  finally
    ccall((:krylov_workspace_free, lib), Cint, (Ptr{Cvoid},), h[])
  end
  ws
end
# NOTE for the maintainer: a production binding keeps the C handle inside the workspace (created once in the
# CgWorkspace(kc) constructor) and attaches A with krylov_b200_share_operator, so in-place solves allocate
# nothing (test/test_allocations.jl:54-57); the sketch above creates it per call for brevity.
Krylov.cg!(ws::CgWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:cg, ws, A, b; kw...)
Krylov.minres!(ws::MinresWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:minres, ws, A, b; kw...)
Krylov.gmres!(ws::GmresWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:gmres, ws, A, b; memory = length(ws.c), kw...)
Krylov.bicgstab!(ws::BicgstabWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:bicgstab, ws, A, b; kw...)
# sibling solvers served by the same library (SURVEY.md 8f-3); every other method keeps running through the k*
# overloads above, one kernel per call
Krylov.cr!(ws::CrWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:cr, ws, A, b; kw...)
Krylov.cgs!(ws::CgsWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:cgs, ws, A, b; kw...)
Krylov.cg_lanczos!(ws::CgLanczosWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:cg_lanczos, ws, A, b; kw...)
Krylov.fom!(ws::FomWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:fom, ws, A, b; memory = length(ws.l), kw...)
Krylov.fgmres!(ws::FgmresWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:fgmres, ws, A, b; memory = length(ws.c), kw...)
Krylov.dqgmres!(ws::DqgmresWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:dqgmres, ws, A, b; memory = length(ws.V), kw...)
Krylov.diom!(ws::DiomWorkspace{T,T,B200Vector{T}}, A::B200CSR{T}, b::B200Vector{T}; kw...) where T = fused_solve!(:diom, ws, A, b; memory = length(ws.V), kw...)

end # module
