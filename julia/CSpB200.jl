# CSpB200.jl -- Julia binding of libcsp_b200.so (include/csp_b200.h) for CoordinateSplittingPTrees.jl.
#
# The reference package stays as it is: World, Box{p}, addpoint!, value, position, leaves ... keep growing and inspecting the
# tree on the host.  This module ADDS batched methods to the reference's own functions for the data-parallel hot path
# (find_leaf_at, fit_quadratic, coefficients_p, modelvalue, minimum / maximum / extrema, possemidef); each flattens the Box graph
# with the reference's preorder iterator (src/tree.jl:393-415) and `ccall`s the library.  Every exported C entry point is bound here.
#
# NOTE: Julia is not installed in the image this library is built and tested in, so this file has not been executed there; the
# ccall signatures are transcriptions of include/csp_b200.h, and the same entry points with the same argument layouts are
# exercised by the Python twin (coordinatesplittingptrees.jl_b200/{_native,device,api,shard}.py), which the test-suite runs.
module CSpB200

using CoordinateSplittingPTrees
using LinearAlgebra: SymTridiagonal
const CSp = CoordinateSplittingPTrees
import CoordinateSplittingPTrees: find_leaf_at, Box, isleaf, isfake, value, position, leaves, boxbounds
import Base: minimum, maximum, extrema

const LIB = get(ENV, "CSP_B200_LIB", joinpath(@__DIR__, "..", "coordinatesplittingptrees.jl_b200", "libcsp_b200.so"))

# ---- status codes (include/csp_b200.h) ---------------------------------------------------------------------------------------
const CSP_OK = 0
const Q_OUTSIDE = 0x01
const FIT_OK, FIT_NO_CHAIN, FIT_CHAIN_ASSERT, FIT_CHAIN_UNDEF = 0x00, 0x01, 0x02, 0x03

last_error() = unsafe_string(ccall((:csp_last_error, LIB), Cstring, ()))
"Library errors become Julia exceptions: CSP_ERR_INVALID -> ArgumentError, everything else ErrorException."
function check(rc::Integer)
    rc == CSP_OK && return nothing
    rc == -1 ? throw(ArgumentError(last_error())) : error(last_error())
end

version() = ccall((:csp_version, LIB), Cint, ())
function device_count()
    c = Ref{Cint}(0); check(ccall((:csp_device_count, LIB), Cint, (Ref{Cint},), c)); Int(c[])
end
set_device(dev::Integer) = check(ccall((:csp_set_device, LIB), Cint, (Cint,), dev))
kernel_launches() = ccall((:csp_kernel_launches, LIB), Int64, ())
timing_enable(on::Bool=true) = check(ccall((:csp_timing_enable, LIB), Cint, (Cint,), on))
function timing_read()
    ms = zeros(Float64, 7); cnt = zeros(Int64, 7)
    check(ccall((:csp_timing_read, LIB), Cint, (Ptr{Float64}, Ptr{Int64}), ms, cnt))
    names = (:locate, :locate_eval_fused, :scan, :scatter, :eval_binned, :eval, :fit)
    Dict(k => (ms[i], cnt[i]) for (i, k) in enumerate(names))
end

# ---- flattening the Box graph --------------------------------------------------------------------------------------------------
struct TreeDesc                      # == csp_tree_desc
    n::Int32; p::Int32; n_nodes::Int32; n_internal::Int32; n_leaves::Int32
    lower::Ptr{Float64}; upper::Ptr{Float64}
    parent::Ptr{Int32}; childindex::Ptr{UInt8}; internal_id::Ptr{Int32}; leaf_id::Ptr{Int32}; val::Ptr{Float64}
    inode::Ptr{Int32}; dims::Ptr{Int32}; xs::Ptr{Float64}; child::Ptr{Int32}; pos::Ptr{Float64}
end

"The csp_tree_desc arrays of `collect(root)` (preorder), plus the box <-> id maps."
struct FlatTree
    n::Int; p::Int
    lower::Vector{Float64}; upper::Vector{Float64}
    parent::Vector{Int32}; childindex::Vector{UInt8}; internal_id::Vector{Int32}; leaf_id::Vector{Int32}; val::Vector{Float64}
    inode::Vector{Int32}; dims::Vector{Int32}; xs::Vector{Float64}; child::Vector{Int32}; pos::Vector{Float64}
    boxes::Vector{Any}         # node id + 1 -> Box
    leafboxes::Vector{Any}     # leaf id + 1 -> Box   (leaves(root), src/tree.jl:423)
    nodeid::IdDict{Any,Int32}  # Box -> node id
end

function flatten(root::Box{p,Float64}) where p
    boxes = collect(root); n = ndims(root); C = 1 << p
    nodeid = IdDict{Any,Int32}(b => Int32(i - 1) for (i, b) in enumerate(boxes))
    nn = length(boxes)
    internal = [b for b in boxes if !isleaf(b)]
    parent = Int32[nodeid[CSp.isroot(b) ? b : b.parent] for b in boxes]
    childindex = UInt8[b.childindex for b in boxes]
    internal_id = fill(Int32(-1), nn); leaf_id = fill(Int32(-1), nn)
    I = 0; L = 0; leafboxes = Any[]
    for (i, b) in enumerate(boxes)
        if !isleaf(b)
            internal_id[i] = I; I += 1
        elseif !isfake(b)                                   # src/tree.jl:430-442
            leaf_id[i] = L; L += 1; push!(leafboxes, b)
        end
    end
    val   = Float64[value(b) for b in boxes]                  # meta walk, src/tree.jl:316-323
    inode = Int32[nodeid[b] for b in internal]
    dims  = Int32[d for b in internal for d in b.split.dims]
    xs    = Float64[x for b in internal for x in b.split.xs]
    child = Int32[nodeid[CSp.getchild(b.split, k)] for b in internal for k in 0:C-1]   # src/tree.jl:325
    pos   = isempty(internal) ? Float64[] : reduce(vcat, (collect(Float64, position(b)) for b in internal))
    w = root.world
    FlatTree(n, p, collect(Float64, w.lower), collect(Float64, w.upper), parent, childindex, internal_id, leaf_id, val,
             inode, dims, xs, child, pos, boxes, leafboxes, nodeid)
end

function with_desc(f, ft::FlatTree)
    GC.@preserve ft begin
        d = TreeDesc(ft.n, ft.p, length(ft.parent), length(ft.inode), length(ft.leafboxes), pointer(ft.lower), pointer(ft.upper),
                     pointer(ft.parent), pointer(ft.childindex), pointer(ft.internal_id), pointer(ft.leaf_id), pointer(ft.val),
                     pointer(ft.inode), pointer(ft.dims), pointer(ft.xs), pointer(ft.child), pointer(ft.pos))
        f(Ref(d))
    end
end

"The flattened tree as one relocatable byte blob (wire / on-disk format): csp_tree_blob_size + csp_tree_pack."
function pack(ft::FlatTree)
    with_desc(ft) do d
        sz = Ref{Csize_t}(0)
        check(ccall((:csp_tree_blob_size, LIB), Cint, (Ref{TreeDesc}, Ref{Csize_t}), d, sz))
        blob = zeros(UInt8, sz[])
        check(ccall((:csp_tree_pack, LIB), Cint, (Ref{TreeDesc}, Ptr{UInt8}, Csize_t), d, blob, sz[]))
        blob
    end
end

# ---- device-resident tree ---------------------------------------------------------------------------------------------------------
mutable struct DeviceTree
    h::Ptr{Cvoid}
    flat::Union{FlatTree,Nothing}    # nothing when built from a blob on a rank that never saw the Box graph
    owned::Bool
end
free!(t::DeviceTree) = (t.owned && t.h != C_NULL && ccall((:csp_tree_free, LIB), Cint, (Ptr{Cvoid},), t.h); t.h = C_NULL; nothing)

"Flatten `root` and upload it (csp_tree_upload).  Call again after `addpoint!`: preorder ids change with every split."
function DeviceTree(root::Box)
    ft = flatten(root)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    with_desc(ft) do d
        check(ccall((:csp_tree_upload, LIB), Cint, (Ref{TreeDesc}, Ref{Ptr{Cvoid}}), d, h))
    end
    finalizer(free!, DeviceTree(h[], ft, true))
end
"Upload a packed blob; `blob` may be a host array or a device pointer with its length (csp_tree_upload_blob)."
function DeviceTree(blob::Vector{UInt8}; flat=nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_tree_upload_blob, LIB), Cint, (Ptr{UInt8}, Csize_t, Ref{Ptr{Cvoid}}), blob, length(blob), h))
    finalizer(free!, DeviceTree(h[], flat, true))
end
function DeviceTree(dblob::Ptr{Cvoid}, bytes::Integer; flat=nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_tree_upload_blob, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), dblob, bytes, h))
    finalizer(free!, DeviceTree(h[], flat, true))
end
"(n, p, n_nodes, n_internal, n_leaves, device, max leaf depth, KiB resident)"
function info(t::DeviceTree)
    v = zeros(Int64, 8); check(ccall((:csp_tree_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), t.h, v)); Tuple(v)
end
nodeid(t::DeviceTree, box) = t.flat.nodeid[box]
leafbox(t::DeviceTree, leaf::Integer) = t.flat.leafboxes[leaf + 1]

# ---- live trees: the device copy grows with the tree (include/csp_b200.h, "incremental residency") ------------------------------------------
"""
A device tree fed the SPLIT LOG instead of whole flattenings: `LiveTree(root)` replays the splits of `root` in creation order and
`update!(lt)` sends only the splits `addpoint!` made since the last call; the preorder flattening is re-derived on the device.
Box ids here are creation-order ids (root 0; the children of the s-th split are 1 + s*2^p + (0:2^p-1)).
"""
mutable struct LiveTree
    tree::DeviceTree
    root::Box
    boxid::IdDict{Any,Int32}   # Box -> creation-order id
    nsplits::Int
end
function LiveTree(root::Box{p,Float64}) where p
    w = root.world; n = ndims(root)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_tree_create_live, LIB), Cint, (Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
                n, p, collect(Float64, w.lower), collect(Float64, w.upper), collect(Float64, position(root)), value(root), h))
    lt = LiveTree(finalizer(free!, DeviceTree(h[], nothing, true)), root, IdDict{Any,Int32}(root => Int32(0)), 0)
    update!(lt)
end
"Append the splits made since the last call.  The reference keeps no log, so they are found by walking the boxes that were leaves."
function update!(lt::LiveTree)
    p = CSp.degree(lt.root); C = 1 << p
    sb = Int32[]; dims = Int32[]; xs = Float64[]; cv = Float64[]
    frontier = [b for b in keys(lt.boxid) if !isleaf(b) && !haskey(lt.boxid, CSp.getchild(b.split, 0))]
    sort!(frontier; by = b -> lt.boxid[b])
    while !isempty(frontier)
        b = popfirst!(frontier)
        push!(sb, lt.boxid[b]); append!(dims, b.split.dims); append!(xs, b.split.xs)
        for k in 0:C-1
            c = CSp.getchild(b.split, k)
            lt.boxid[c] = Int32(1 + (lt.nsplits + length(sb) - 1) * C + k)
            push!(cv, value(c))
            isleaf(c) || push!(frontier, c)
        end
    end
    isempty(sb) || check(ccall((:csp_tree_append_splits, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                               lt.tree.h, length(sb), sb, dims, xs, cv))
    lt.nsplits += length(sb)
    lt
end
"(node_of_box, box_of_node): creation-order box ids <-> the current preorder node ids."
function live_maps(lt::LiveTree)
    nn = info(lt.tree)[3]; a = Vector{Int32}(undef, nn); b = Vector{Int32}(undef, nn)
    check(ccall((:csp_tree_live_maps, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), lt.tree.h, a, b))
    (a, b)
end

"Diagnostic: one of the arrays derived on the device at upload, as raw bytes (csp_tree_export_derived)."
function export_derived(t::DeviceTree, which::Integer)
    sz = Ref{Csize_t}(0)
    check(ccall((:csp_tree_export_derived, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Ref{Csize_t}), t.h, which, C_NULL, 0, sz))
    out = zeros(UInt8, sz[])
    sz[] > 0 && check(ccall((:csp_tree_export_derived, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}, Csize_t, Ref{Csize_t}), t.h, which, out, sz[], sz))
    out
end

# ---- K1: find_leaf_at ---------------------------------------------------------------------------------------------------------------
function _throw_outside(X, status)
    bad = findfirst(!iszero, status)
    bad === nothing || error("$(X[:, bad]) not within the bounds of the box (query $bad)")     # src/tree.jl:359
end

"Batched find_leaf_at (src/tree.jl:351-388): X is n x m, each column a query.  Returns leaf ids (ranks in leaves(root))."
function find_leaf_ids(t::DeviceTree, X::Matrix{Float64}; from=nothing)
    n = info(t)[1]
    size(X, 1) == n || throw(DimensionMismatch("tree has $n dimensions, got $(size(X, 1))"))              # src/tree.jl:353
    m = size(X, 2); leaf = Vector{Int32}(undef, m); status = Vector{UInt8}(undef, m)
    if from === nothing
        check(ccall((:csp_find_leaf, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{UInt8}), t.h, X, m, leaf, status))
    else   # any box as the starting point, general boxbounds check (src/tree.jl:146-171, 357-360)
        check(ccall((:csp_find_leaf_from, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{UInt8}),
                    t.h, nodeid(t, from), X, m, leaf, status))
    end
    _throw_outside(X, status)
    leaf
end
"find_leaf_at(tree, X) / find_leaf_at(tree, box, X): the leaf boxes containing the columns of X."
find_leaf_at(t::DeviceTree, X::Matrix{Float64}) = [leafbox(t, l) for l in find_leaf_ids(t, X)]
find_leaf_at(t::DeviceTree, box::Box, X::Matrix{Float64}) = [leafbox(t, l) for l in find_leaf_ids(t, X; from=box)]
find_leaf_at(t::DeviceTree, box::Box, x::AbstractVector) = find_leaf_at(t, box, reshape(collect(Float64, x), :, 1))[1]

"Resident queries: dX / dleaf / dstatus are device pointers (e.g. from CUDA.jl `pointer(CuArray)`), `stream` a CUstream."
find_leaf_dev!(t::DeviceTree, dX::Ptr{Cvoid}, m::Integer, dleaf::Ptr{Cvoid}, dstatus::Ptr{Cvoid}, stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_find_leaf_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), t.h, dX, m, dleaf, dstatus, stream))
find_leaf_from_dev!(t::DeviceTree, box::Box, dX::Ptr{Cvoid}, m::Integer, dleaf::Ptr{Cvoid}, dstatus::Ptr{Cvoid}, stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_find_leaf_from_dev, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                t.h, nodeid(t, box), dX, m, dleaf, dstatus, stream))

"boxbounds(box) (src/tree.jl:136-171) computed on the device."
function boxbounds(t::DeviceTree, box::Box)
    n = info(t)[1]; lo = zeros(n); hi = zeros(n)
    check(ccall((:csp_boxbounds, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), t.h, nodeid(t, box), lo, hi))
    collect(zip(lo, hi))
end

"extrema over leaves(box) (src/CoordinateSplittingPTrees.jl:64-68): (minimum leaf, maximum leaf), first of equal leaves."
function extrema(t::DeviceTree, box::Box=t.flat.boxes[1])
    lmin = Ref{Int32}(0); lmax = Ref{Int32}(0); vmin = Ref{Float64}(0); vmax = Ref{Float64}(0)
    check(ccall((:csp_tree_extrema, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Float64}, Ref{Int32}, Ref{Float64}),
                t.h, nodeid(t, box), lmin, vmin, lmax, vmax))
    (leafbox(t, lmin[]), leafbox(t, lmax[]))
end
minimum(t::DeviceTree, box::Box=t.flat.boxes[1]) = extrema(t, box)[1]
maximum(t::DeviceTree, box::Box=t.flat.boxes[1]) = extrema(t, box)[2]

# ---- K2: fits -----------------------------------------------------------------------------------------------------------------------
function _raise_fit(st::UInt8, box)
    st == FIT_NO_CHAIN && error("no chain found at $box")                                             # src/cs2.jl:167
    st == FIT_CHAIN_ASSERT && throw(AssertionError("!(wassplit[d1]) && (d2 > n || !(wassplit[d2]))"))   # src/cs2.jl:186-187
    st == FIT_CHAIN_UNDEF && throw(UndefRefError())                                                   # src/cs2.jl:184
end

"fit_quadratic_native(box; skip) for a vector of boxes (src/cs2.jl:115-121): (c, g, Q, b, y, prec, firstmatch) per box."
function fit_quadratic_native(t::DeviceTree, boxes::Vector; skip::Int=0, throw_on_failure::Bool=true)
    n = info(t)[1]; nb = length(boxes)
    ids = Int32[nodeid(t, b) for b in boxes]
    c = Vector{Float64}(undef, nb); g = Matrix{Float64}(undef, n, nb); Q = Array{Float64,3}(undef, n, n, nb)
    b = Matrix{Float64}(undef, n, nb); gn = similar(g); y = similar(g)
    prec = Array{UInt8,3}(undef, n, n, nb); fm = Matrix{Int32}(undef, n, nb); st = Vector{UInt8}(undef, nb)
    check(ccall((:csp_fit_boxes, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int32}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int32}, Ptr{UInt8}),
                t.h, ids, nb, skip, c, g, Q, b, gn, y, prec, fm, st))
    map(1:nb) do i
        throw_on_failure && _raise_fit(st[i], boxes[i])
        # Q[:, :, i] already is the SymmetricArray .data layout (src/types.jl:432-454): values at [i,j], i <= j, NaN below
        (c = c[i], g = g[:, i], gnative = gn[:, i], Q = CSp.SymmetricArray(Q[:, :, i]), b = b[:, i], y = y[:, i],
         prec = prec[:, :, i] .!= 0, firstmatch = Int.(fm[:, i]), status = st[i])
    end
end
"fit_quadratic(box; skip) (src/cs2.jl:89-93) for a vector of boxes: (c, g, Q, b) per box."
fit_quadratic(t::DeviceTree, boxes::Vector; skip::Int=0) = [(r.c, r.g, r.Q, r.b) for r in fit_quadratic_native(t, boxes; skip=skip)]
fit_quadratic(t::DeviceTree, box::Box; skip::Int=0) = fit_quadratic(t, [box]; skip=skip)[1]

"coefficients_p(box; skip) (src/polynomials.jl:127-230) for a vector of boxes; `callback(count)` receives the number of times the
reference would have invoked its per-child callback (2^p per split that supplied a coefficient)."
function coefficients_p(t::DeviceTree, boxes::Vector; skip::Int=0, callback=nothing)
    n, p = info(t)[1], info(t)[2]; nb = length(boxes)
    ids = Int32[nodeid(t, b) for b in boxes]
    Cp = Array{Float64}(undef, ntuple(_ -> n, p)..., nb)
    visited = Vector{Int64}(undef, nb); used = Vector{Int64}(undef, nb)
    check(ccall((:csp_coefficients_p, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int64, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                t.h, ids, nb, skip, Cp, visited, used))
    callback === nothing || foreach(u -> callback(u << p), used)
    [CSp.SymmetricArray(copy(selectdim(Cp, p + 1, i))) for i in 1:nb]
end
coefficients_p(t::DeviceTree, box::Box; kw...) = coefficients_p(t, [box]; kw...)[1]

"coefficients_p! into a sparse symmetric output (src/polynomials.jl:161-230, 244-271): `pattern` lists the stored above-diagonal
positions (i, j), i < j; for a SymTridiagonal output use `tridiagonal(t, boxes)`."
function coefficients_p_sparse(t::DeviceTree, boxes::Vector, pattern::Vector{Tuple{Int,Int}}; skip::Int=0, callback=nothing)
    nb = length(boxes); np = length(pattern)
    ids = Int32[nodeid(t, b) for b in boxes]
    pi = Int32[min(a, b) for (a, b) in pattern]; pj = Int32[max(a, b) for (a, b) in pattern]
    vals = Matrix{Float64}(undef, np, nb); visited = Vector{Int64}(undef, nb); used = Vector{Int64}(undef, nb)
    check(ccall((:csp_coefficients_p_sparse, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int32}, Int64, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                t.h, ids, nb, skip, np, pi, pj, vals, visited, used))
    callback === nothing || foreach(u -> callback(u << 2), used)
    vals
end
function tridiagonal(t::DeviceTree, boxes::Vector; kw...)
    n = info(t)[1]
    ev = coefficients_p_sparse(t, boxes, [(i, i + 1) for i in 1:n-1]; kw...)
    [SymTridiagonal(fill(NaN, n), ev[:, i]) for i in 1:length(boxes)]
end

"Q = coefficients_p(box) followed by fit_quadratic_gdiag!(Q, box) (src/cs2.jl:332-384): the chain-free fit, (c, g, Q, b) per box."
function fit_quadratic_gdiag(t::DeviceTree, boxes::Vector; skip::Int=0)
    n = info(t)[1]; nb = length(boxes)
    ids = Int32[nodeid(t, b) for b in boxes]
    c = Vector{Float64}(undef, nb); g = Matrix{Float64}(undef, n, nb); Q = Array{Float64,3}(undef, n, n, nb); b = similar(g)
    check(ccall((:csp_fit_gdiag, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                t.h, ids, nb, skip, c, g, Q, b))
    [(c[i], g[:, i], CSp.SymmetricArray(Q[:, :, i]), b[:, i]) for i in 1:nb]
end

"polynomial_full (README.md:98-99; not defined in the reference's src/): p = 2 -> fit_quadratic; p = 1 -> the degree-1 model
(c, g, b) with c = value(box), g = coefficients_p(box), b = position(box)."
polynomial_full(t::DeviceTree, box::Box{2}; skip=0) = fit_quadratic(t, box; skip=skip)
polynomial_full(t::DeviceTree, box::Box{1}; skip=0) = (value(box), coefficients_p(t, box; skip=skip).data, position(box))
"polynomial_minimal (README only, no reference definition): the chain-only part of the CS2 model, Q's diagonal left NaN."
function polynomial_minimal(t::DeviceTree, box::Box{2}; skip=0)
    r = fit_quadratic_native(t, [box]; skip=skip)[1]
    Qd = copy(r.Q.data); for i in 1:size(Qd, 1); Qd[i, i] = NaN; end
    (r.c, r.gnative, CSp.SymmetricArray(Qd), r.b, r.y, r.prec)
end

# ---- per-leaf models resident on the device ----------------------------------------------------------------------------------------
mutable struct LeafModels
    h::Ptr{Cvoid}
    owned::Bool
end
free!(m::LeafModels) = (m.owned && m.h != C_NULL && ccall((:csp_models_free, LIB), Cint, (Ptr{Cvoid},), m.h); m.h = C_NULL; nothing)
"(n, n_models, degree, device)"
function info(m::LeafModels)
    v = zeros(Int64, 4); check(ccall((:csp_models_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), m.h, v)); Tuple(v)
end
"Fit every leaf and keep (c, g, Q, b) on the device (csp_models_fit)."
function fit_leaves(t::DeviceTree; skip::Int=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_models_fit, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), t.h, skip, h))
    finalizer(free!, LeafModels(h[], true))
end
function alloc_models(t::DeviceTree)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_models_alloc, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), t.h, h))
    finalizer(free!, LeafModels(h[], true))
end
refit!(t::DeviceTree, m::LeafModels; skip::Int=0, stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_models_refit, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}), t.h, skip, m.h, stream))
"Re-fit only the leaves lo:hi-1 (0-based half-open range of leaf ids): the per-GPU piece of a leaf-sharded fit."
refit_range!(t::DeviceTree, m::LeafModels, lo::Integer, hi::Integer; skip::Int=0, stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_models_refit_range, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}), t.h, skip, m.h, lo, hi, stream))
"Caller-supplied models: c[nl], g[n, nl], Q[n, n, nl] (.data layout; `nothing` for degree-1 models), b[n, nl]."
function upload_models(c::Vector{Float64}, g::Matrix{Float64}, Q::Union{Array{Float64,3},Nothing}, b::Matrix{Float64})
    n, nl = size(g); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_models_upload, LIB), Cint, (Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                n, nl, c, g, Q === nothing ? C_NULL : Q, b, h))
    finalizer(free!, LeafModels(h[], true))
end
function download(m::LeafModels)
    n, nl, degree, _ = info(m)
    c = Vector{Float64}(undef, nl); g = Matrix{Float64}(undef, n, nl); b = similar(g); st = Vector{UInt8}(undef, nl)
    Q = degree == 2 ? Array{Float64,3}(undef, n, n, nl) : nothing
    check(ccall((:csp_models_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                m.h, c, g, Q === nothing ? C_NULL : Q, b, st))
    (c = c, g = g, Q = Q, b = b, status = st)
end
function download_rows(m::LeafModels, leaf_ids::Vector{Int32})
    n, _, degree, _ = info(m); nb = length(leaf_ids)
    c = Vector{Float64}(undef, nb); g = Matrix{Float64}(undef, n, nb); b = similar(g); st = Vector{UInt8}(undef, nb)
    Q = degree == 2 ? Array{Float64,3}(undef, n, n, nb) : nothing
    check(ccall((:csp_models_download_rows, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                m.h, leaf_ids, nb, c, g, Q === nothing ? C_NULL : Q, b, st))
    (c = c, g = g, Q = Q, b = b, status = st)
end
"possemidef (src/posdef.jl:41-68) of every resident model's Q, in place on the device."
possemidef!(m::LeafModels; stream::Ptr{Cvoid}=C_NULL) = check(ccall((:csp_models_possemidef, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), m.h, stream))

# ---- K3: evaluation -------------------------------------------------------------------------------------------------------------------
"modelvalue(c, g, Q, b, x) (src/cs2.jl:155-159) of model leaf[i] at X[:, i]; a leaf id outside the models gives NaN."
function modelvalue(m::LeafModels, X::Matrix{Float64}, leaf::Vector{Int32})
    val = Vector{Float64}(undef, size(X, 2))
    check(ccall((:csp_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int64, Ptr{Float64}), m.h, X, leaf, size(X, 2), val))
    val
end
eval_dev!(m::LeafModels, dX::Ptr{Cvoid}, dleaf::Ptr{Cvoid}, n::Integer, dval::Ptr{Cvoid}, stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_eval_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), m.h, dX, dleaf, n, dval, stream))
"find_leaf_at followed by modelvalue of that leaf's model, fused: (leaf ids, values)."
function locate_eval(t::DeviceTree, m::LeafModels, X::Matrix{Float64})
    nq = size(X, 2); leaf = Vector{Int32}(undef, nq); val = Vector{Float64}(undef, nq); status = Vector{UInt8}(undef, nq)
    check(ccall((:csp_locate_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}),
                t.h, m.h, X, nq, leaf, val, status))
    _throw_outside(X, status)
    (leaf, val)
end
locate_eval_dev!(t::DeviceTree, m::LeafModels, dX::Ptr{Cvoid}, nq::Integer, dleaf::Ptr{Cvoid}, dval::Ptr{Cvoid}, dstatus::Ptr{Cvoid},
                 stream::Ptr{Cvoid}=C_NULL) =
    check(ccall((:csp_locate_eval_dev, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                t.h, m.h, dX, nq, dleaf, dval, dstatus, stream))
"modelvalue(c, g, Q, b, y, prec, x) for the CS2-native model (src/cs2.jl:142-153), columns of X."
function modelvalue_native(c::Float64, g::Vector{Float64}, Q::Matrix{Float64}, b::Vector{Float64}, y::Vector{Float64}, prec::AbstractMatrix{Bool},
                           X::Matrix{Float64})
    n, m = size(X); val = Vector{Float64}(undef, m); P = UInt8.(prec)
    check(ccall((:csp_modelvalue_native, LIB), Cint,
                (Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Int64, Ptr{Float64}),
                n, c, g, Q, b, y, P, X, m, val))
    val
end

# ---- multi-GPU: one Julia process drives every GPU of the box (include/csp_b200.h, "multi-GPU") ------------------------------------------
"csp_init: GPUs devs (default 0:ndev-1) become ranks 0:ndev-1 of one NCCL clique."
init(ndev::Integer, devs::Union{Vector{Int32},Nothing}=nothing) =
    check(ccall((:csp_init, LIB), Cint, (Cint, Ptr{Int32}), ndev, devs === nothing ? C_NULL : devs))
shutdown() = check(ccall((:csp_shutdown, LIB), Cint, ()))
"(world size, local ranks, first local rank, NCCL version code)"
function comm_info()
    v = zeros(Int64, 4); check(ccall((:csp_comm_info, LIB), Cint, (Ptr{Int64},), v)); Tuple(v)
end
"One-process-per-GPU mode (e.g. under MPI.jl): rank 0 creates the id, every rank joins."
function comm_unique_id()
    id = zeros(UInt8, 128); check(ccall((:csp_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id
end
init_rank(world::Integer, rank::Integer, device::Integer, id::Vector{UInt8}) =
    check(ccall((:csp_init_rank, LIB), Cint, (Cint, Cint, Cint, Ptr{UInt8}), world, rank, device, id))

mutable struct ShardedTree
    h::Ptr{Cvoid}
    flat::Union{FlatTree,Nothing}
end
mutable struct ShardedModels
    h::Ptr{Cvoid}
end
free!(s::ShardedTree) = (s.h != C_NULL && ccall((:csp_shard_tree_free, LIB), Cint, (Ptr{Cvoid},), s.h); s.h = C_NULL; nothing)
free!(s::ShardedModels) = (s.h != C_NULL && ccall((:csp_shard_models_free, LIB), Cint, (Ptr{Cvoid},), s.h); s.h = C_NULL; nothing)

"C1: flatten, pack, broadcast once from rank `root` and upload on every GPU (csp_shard_tree_broadcast)."
function ShardedTree(rootbox::Box; root::Integer=0)
    ft = flatten(rootbox); blob = pack(ft); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_shard_tree_broadcast, LIB), Cint, (Ptr{UInt8}, Csize_t, Cint, Ref{Ptr{Cvoid}}), blob, length(blob), root, h))
    finalizer(free!, ShardedTree(h[], ft))
end
"The replica on local GPU i (0-based), as a borrowed DeviceTree."
function local_tree(s::ShardedTree, i::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_shard_tree_get, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), s.h, i, h))
    DeviceTree(h[], s.flat, false)
end
"Leaf fits sharded by leaf-id range over the GPUs, fitted records all-gathered (csp_shard_fit).  Pass `models` to re-fit in place."
function fit_leaves(s::ShardedTree; skip::Int=0, models::Union{ShardedModels,Nothing}=nothing)
    h = Ref{Ptr{Cvoid}}(models === nothing ? C_NULL : models.h)
    check(ccall((:csp_shard_fit, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}), s.h, skip, h, C_NULL))
    models === nothing ? finalizer(free!, ShardedModels(h[])) : models
end
function local_models(sm::ShardedModels, i::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_shard_models_get, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), sm.h, i, h))
    LeafModels(h[], false)
end
"Host buffers: the batch is split over the GPUs (one copy / compute pipeline per GPU) and gathered in the caller's arrays."
function locate_eval(s::ShardedTree, sm::ShardedModels, X::Matrix{Float64})
    nq = size(X, 2); leaf = Vector{Int32}(undef, nq); val = Vector{Float64}(undef, nq); status = Vector{UInt8}(undef, nq)
    check(ccall((:csp_shard_locate_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}),
                s.h, sm.h, X, nq, leaf, val, status))
    _throw_outside(X, status)
    (leaf, val)
end
"Resident shards (device pointers per local GPU); with gather_root >= 0 the (leaf, value) pairs are gathered on that rank (C2)."
function locate_eval_dev!(s::ShardedTree, sm::ShardedModels, dX::Vector{Ptr{Cvoid}}, counts::Vector{Int64}, dleaf::Vector{Ptr{Cvoid}},
                          dval::Vector{Ptr{Cvoid}}, dstatus::Vector{Ptr{Cvoid}}; gather_root::Integer=-1, dleaf_all::Ptr{Cvoid}=C_NULL,
                          dval_all::Ptr{Cvoid}=C_NULL, gather_base::Integer=0, chunk::Integer=0,
                          streams::Union{Vector{Ptr{Cvoid}},Nothing}=nothing)
    check(ccall((:csp_shard_locate_eval_dev, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}, Ptr{Cvoid},
                 Int64, Int64, Ptr{Ptr{Cvoid}}),
                s.h, sm.h, dX, counts, dleaf, dval, dstatus, gather_root, dleaf_all, dval_all, gather_base, chunk,
                streams === nothing ? C_NULL : streams))
end
gather_dev!(dleaf::Vector{Ptr{Cvoid}}, dval::Vector{Ptr{Cvoid}}, counts::Vector{Int64}, gather_root::Integer, dleaf_all::Ptr{Cvoid},
            dval_all::Ptr{Cvoid}; gather_base::Integer=0, streams::Union{Vector{Ptr{Cvoid}},Nothing}=nothing) =
    check(ccall((:csp_shard_gather_dev, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}),
                dleaf, dval, counts, gather_root, dleaf_all, dval_all, gather_base, streams === nothing ? C_NULL : streams))
"Library-owned gather buffers on rank `root` (copy-engine gather: peer DMA over NVLink); returns (dleaf_all, dval_all) device pointers."
function gather_alloc(capacity::Integer; root::Integer=0)
    pl = Ref{Ptr{Cvoid}}(C_NULL); pv = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csp_shard_gather_alloc, LIB), Cint, (Int64, Cint, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}), capacity, root, pl, pv))
    (pl[], pv[])
end
gather_free() = check(ccall((:csp_shard_gather_free, LIB), Cint, ()))
"Let consecutive resident calls overlap the previous call's last gather step; `join()` then waits for all gathers issued so far."
defer_join(on::Bool=true) = check(ccall((:csp_shard_defer_join, LIB), Cint, (Cint,), on))
join(streams::Union{Vector{Ptr{Cvoid}},Nothing}=nothing) = check(ccall((:csp_shard_join, LIB), Cint, (Ptr{Ptr{Cvoid}},), streams === nothing ? C_NULL : streams))
synchronize() = check(ccall((:csp_shard_synchronize, LIB), Cint, ()))
"(blob h2d, tree broadcast, upload, fit kernels, record all-gather, result gather) in milliseconds"
function shard_timing()
    ms = zeros(Float64, 6); check(ccall((:csp_shard_timing, LIB), Cint, (Ptr{Float64},), ms)); Tuple(ms)
end

end # module
