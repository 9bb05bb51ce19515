# SpaSM_b200.jl — Julia host for libspasm_b200.so (UNEXECUTED here: no Julia in the build image).
#
# Two ways to use the CUDA library from Julia:
#
#  (1) drop-in: SpaSM.jl binds everything through ONE constant (src/SpaSM.jl:14).  Point it at this
#      library and every `@ccall spasm_lib.spasm_*` of the hot path runs on the B200:
#
#          const spasm_lib = get(ENV, "SPASM_B200_LIB", Spasm_jll.spasm)
#
#  (2) this stand-alone module: the minimal API of the north-star path (CSR, echelonize, rank,
#      kernel, solve) with explicit NULL checks (the library returns NULL/false instead of aborting
#      when CUDA fails; SpaSM.jl's `LU(ptr)` would fault on a NULL).
module SpaSM_b200

using SparseArrays, Libdl

const lib = get(ENV, "SPASM_B200_LIB", joinpath(@__DIR__, "..", "libspasm_b200.so"))

struct Field            # include/spasm_b200.h: struct spasm_field_struct (32 B)
    p::Int64; halfp::Int64; mhalfp::Int64; dinvp::Float64
end
struct CSRHeader        # struct spasm_csr (72 B)
    nzmax::Int64; n::Int32; m::Int32
    p::Ptr{Int64}; j::Ptr{Int32}; x::Ptr{Int32}
    field::Field
end
struct LUHeader         # struct spasm_lu (48 B)
    r::Int32; complete::UInt8
    L::Ptr{CSRHeader}; U::Ptr{CSRHeader}; qinv::Ptr{Int32}; p::Ptr{Int32}; Ltmp::Ptr{Cvoid}
end
mutable struct Opts     # struct echelonize_opts (64 B)
    enable_greedy_pivot_search::Bool; enable_tall_and_skinny::Bool; enable_dense::Bool; enable_GPLU::Bool
    L::Bool; complete::Bool
    min_pivot_proportion::Float64; max_round::Int32; sparsity_threshold::Float64
    dense_block_size::Int64; low_rank_ratio::Float64; tall_and_skinny_ratio::Float64; low_rank_start_weight::Float64
    Opts() = (o = new(); ccall((:spasm_echelonize_init_opts, lib), Cvoid, (Ref{Opts},), o); o)
end

mutable struct CSR
    ptr::Ptr{CSRHeader}
    function CSR(ptr::Ptr{CSRHeader}; own = true)
        ptr == C_NULL && error("libspasm_b200 returned NULL (see stderr)")
        A = new(ptr)
        own && finalizer(a -> ccall((:spasm_csr_free, lib), Cvoid, (Ptr{CSRHeader},), a.ptr), A)
        A
    end
end
header(A::CSR) = unsafe_load(A.ptr)
Base.size(A::CSR) = (h = header(A); (Int(h.n), Int(h.m)))
SparseArrays.nnz(A::CSR) = ccall((:spasm_nnz, lib), Int64, (Ptr{CSRHeader},), A.ptr)

"A column of `M` becomes a SpaSM row (same convention as SpaSM.jl, src/SpaSM.jl:941-968)."
function CSR(M::SparseMatrixCSC{<:Integer}, prime::Integer = 42013)
    rows, cols = size(M)
    ptr = ccall((:spasm_csr_alloc, lib), Ptr{CSRHeader}, (Int32, Int32, Int64, Int64, Bool), cols, rows, nnz(M), prime, true)
    A = CSR(ptr); h = header(A); k = 0
    for c = 1:cols
        unsafe_store!(h.p, k, c)
        for e = M.colptr[c]:M.colptr[c+1]-1
            v = mod(M.nzval[e], prime); 2v > prime && (v -= prime)
            iszero(v) && continue
            k += 1
            unsafe_store!(h.j, Int32(M.rowval[e] - 1), k); unsafe_store!(h.x, Int32(v), k)
        end
    end
    unsafe_store!(h.p, k, cols + 1)
    A
end

"Columns of the result are the SpaSM rows, each sorted (src/SpaSM.jl:1011-1023)."
function SparseArrays.sparse(A::CSR)
    h = header(A); nz = unsafe_load(h.p, h.n + 1)
    I = Int[]; J = Int[]; V = Int32[]
    for i = 1:h.n, e = unsafe_load(h.p, i)+1:unsafe_load(h.p, i + 1)
        push!(I, unsafe_load(h.j, e) + 1); push!(J, i); push!(V, unsafe_load(h.x, e))
    end
    sparse(I, J, V, Int(h.m), Int(h.n))
end

mutable struct LU
    ptr::Ptr{LUHeader}
    function LU(ptr::Ptr{LUHeader})
        ptr == C_NULL && error("spasm_echelonize failed (see stderr)")
        f = new(ptr)
        finalizer(x -> ccall((:spasm_lu_free, lib), Cvoid, (Ptr{LUHeader},), x.ptr), f)
        f
    end
end
rank(f::LU) = Int(unsafe_load(f.ptr).r)

function echelonize(A::CSR; kwargs...)
    o = Opts()
    for (k, v) in kwargs; setproperty!(o, k, v); end
    LU(ccall((:spasm_echelonize, lib), Ptr{LUHeader}, (Ptr{CSRHeader}, Ref{Opts}), A.ptr, o))
end
rank(A::CSR; kw...) = rank(echelonize(A; kw...))
kernel(f::LU) = CSR(ccall((:spasm_kernel, lib), Ptr{CSRHeader}, (Ptr{LUHeader},), f.ptr))
kernel(A::CSR; kw...) = kernel(echelonize(A; kw...))
transpose(A::CSR) = CSR(ccall((:spasm_transpose, lib), Ptr{CSRHeader}, (Ptr{CSRHeader},), A.ptr))

"x with x*A == b, or nothing.  Needs echelonize(A; L=true).  x has one entry per row of A."
function solve(f::LU, b::Vector{Int32})
    h = unsafe_load(f.ptr); h.L == C_NULL && error("M.L is null")
    x = zeros(Int32, unsafe_load(h.L).n)
    ok = ccall((:spasm_solve, lib), Bool, (Ptr{LUHeader}, Ptr{Int32}, Ptr{Int32}), f.ptr, b, x)
    ok ? x : nothing
end

# ---- several GPUs of one box: one Julia process per GPU (include/spasm_b200_ext.h).  Rank 0 creates the 128-byte NCCL id,
# the host framework (MPI.jl / Distributed) ships it to the other ranks, every rank calls dist_init!.  From then on
# `echelonize` is a collective call (every rank, same matrix and options): the dense panels and the non-pivotal rows of the
# Schur complements are split over the ranks.  rank(f) and qinv are valid everywhere; kernel / solve need a complete factor
# (rank 0's by default) and raise on a partial one.
nccl_unique_id() = (id = zeros(UInt8, 128); ccall((:spasm_b200_nccl_unique_id, lib), Cint, (Ptr{UInt8},), id) == 0 || error("NCCL id"); id)
dist_init!(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    ccall((:spasm_b200_dist_init, lib), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id) == 0 || error("spasm_b200_dist_init failed (see stderr)")
dist_finalize!() = ccall((:spasm_b200_dist_finalize, lib), Cvoid, ())
"every rank keeps (and downloads) the rows of U of the dense panels it owns instead of rank 0 holding all of them"
shard_factor!(on::Bool) = ccall((:spasm_b200_dist_shard_factor, lib), Cvoid, (Cint,), on)
"kernel / rref / gesv become collective calls whose rows are split over the ranks (every rank gets the complete result)"
shard_rows!(on::Bool) = ccall((:spasm_b200_dist_shard_rows, lib), Cvoid, (Cint,), on)
"memory policy: keep the device / pinned caches between calls (a loop of same-shaped calls), give them back with trim!()"
set_cache!(on::Bool) = ccall((:spasm_b200_set_cache, lib), Cvoid, (Cint,), on)
trim!() = ccall((:spasm_b200_trim, lib), Cvoid, ())

end # module
