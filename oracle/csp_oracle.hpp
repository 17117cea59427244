// csp_oracle.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// A line-by-line C++ restatement of the hot-path functions of
// timholy/CoordinateSplittingPTrees.jl (pure Julia; `julia` is not installed in
// this image, so the reference itself cannot be executed here).  It keeps the
// reference's data structure (a heap-allocated pointer tree in which nothing is
// stored redundantly, so every position / value lookup walks towards the root)
// and its control flow, including its quirks, so that it can serve as
//   (1) the parity checker for the CUDA path (tests/, __graft_entry__.smoke()),
//   (2) the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`).
// Nothing in the product path may include, link or call this file.
//
// Parity pinning: checked against every hand-written golden of the reference's
// test/runtests.jl for this path (geometry, find_leaf_at, preorder, split
// orders, chains, Qcoef identities, full CS2 models) in tests/test_oracle_*.py.
// The reference's randomised tests are unseeded, so they pin properties versus
// analytic truth, not bits; model *evaluation* order (BLAS dot / generic matmul)
// is not pinned by the reference (tolerance-level parity only).
//
// All `file:line` citations are relative to /root/reference/.
// Dimensions are 1-based in the API (as in Julia); storage is 0-based.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cspo {

// Julia exception kinds that cross the C boundary as error codes.
enum ErrKind { ERR_ERROR = 1, ERR_DIMMISMATCH = 2, ERR_ASSERT = 3, ERR_UNDEFREF = 4, ERR_ARGUMENT = 5, ERR_BOUNDS = 6 };
struct JlError : std::runtime_error {
    int kind;
    JlError(int k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

using Objective = std::function<double(const double* y, int n)>;

// src/types.jl:7-23 (World).  P is restricted to plain reals (baseposition(x::Real)=x, :72).
struct World {
    std::vector<double> lower, upper, position;
    double meta;
};
inline World* make_world(const std::vector<double>& lo, const std::vector<double>& up,
                         const std::vector<double>& pos, double meta) {
    size_t N = lo.size();
    if (!(up.size() == N && pos.size() == N)) throw JlError(ERR_DIMMISMATCH, "lower, upper, and position must have equal length");
    for (size_t i = 0; i < N; ++i)  // world_validate, types.jl:78-81
        if (!(lo[i] <= pos[i] && pos[i] <= up[i])) throw JlError(ERR_ARGUMENT, "position is not within bounds");
    return new World{lo, up, pos, meta};
}

struct Box;
// src/types.jl:155-173 (Children, Split)
struct Split {
    std::vector<int> dims;       // length p, 1-based; > n means fictive
    std::vector<double> xs;      // length p
    Box* self;
    std::vector<Box*> others;    // length L = 2^p - 1  (others.children)
    std::vector<double> metas;   // length L            (others.metas)
};
// src/types.jl:187-192 (Box).  `split == nullptr`  <=>  leaf (`!isdefined(box, :split)`, :312)
struct Box {
    World* world;
    Box* parent;
    uint8_t childindex;
    Split* split;
    int p;
    // bookkeeping that is NOT in the reference (filled by index_tree()):
    int preorder_id = -1, leaf_rank = -1;
};

inline bool isroot(const Box* b) { return b->parent == b; }                       // types.jl:306
inline bool isleaf(const Box* b) { return b->split == nullptr; }                  // types.jl:312
inline bool isself(const Box* b) { return !isroot(b) && b->parent->split->self == b; }  // types.jl:313
inline int ndims(const Box* b) { return (int)b->world->lower.size(); }            // types.jl:317
inline int degree(const Box* b) { return b->p; }                                  // types.jl:318
inline int maxchildren(const Box* b) { return 1 << b->p; }                        // types.jl:322-324

inline Box* make_root(World* w, int p) {  // types.jl:194-201
    if (p > 8) throw JlError(ERR_ERROR, "degree must be less than or equal to 8");
    Box* r = new Box{w, nullptr, 0, nullptr, p};
    r->parent = r;
    return r;
}

inline int findfirst_dim(int d, const std::vector<int>& dims) {  // findfirst(isequal(d), split.dims): 1-based, 0 = nothing
    for (size_t i = 0; i < dims.size(); ++i)
        if (dims[i] == d) return (int)i + 1;
    return 0;
}

// src/tree.jl:78-93
inline double position(const Box* box, int d) {
    if (d < 1 || d > ndims(box)) throw JlError(ERR_BOUNDS, "BoundsError: dimension out of range");
    double dflt = box->world->position[d - 1];
    while (!isroot(box)) {
        const Box* p = box->parent;
        unsigned childindex = box->childindex;
        if (!isself(box)) {
            const Split* split = p->split;
            int i = findfirst_dim(d, split->dims);
            if (i != 0 && ((childindex >> (i - 1)) & 1u) != 0u) return split->xs[i - 1];
        }
        box = p;
    }
    return dflt;
}

// src/tree.jl:103-129
inline void position_fill(std::vector<double>& x, std::vector<char>& filled, const Box* box) {
    int N = ndims(box);
    x.resize(N);
    filled.assign(N, 0);
    for (int i = 0; i < N; ++i) x[i] = box->world->position[i];
    int nfilled = 0;
    while (!isroot(box) && nfilled < N) {
        const Box* p = box->parent;
        const Split* split = p->split;
        unsigned childindex = box->childindex;
        for (int i = 1; i <= degree(box); ++i) {
            int sd = split->dims[i - 1];
            double xo = split->xs[i - 1];
            if (sd > N) continue;  // fictive dimension
            if (((childindex >> (i - 1)) & 1u) != 0u && !filled[sd - 1]) {
                x[sd - 1] = xo;
                filled[sd - 1] = 1;
                nfilled += 1;
            }
        }
        box = p;
    }
}
inline std::vector<double> position(const Box* box) {  // tree.jl:71
    std::vector<double> x;
    std::vector<char> f;
    position_fill(x, f, box);
    return x;
}

// src/tree.jl:146-171
inline std::pair<double, double> boxbounds(const Box* box, int d) {
    if (d < 1 || d > ndims(box)) throw JlError(ERR_BOUNDS, "BoundsError: dimension out of range");
    double lower = box->world->lower[d - 1], upper = box->world->upper[d - 1];
    const Box* box0 = box;
    bool lfilled = false, ufilled = false;
    while (!isroot(box) && !(lfilled & ufilled)) {
        const Box* p = box->parent;
        const Split* split = p->split;
        int i = findfirst_dim(d, split->dims);
        if (i != 0) {
            double xs = position(split->self, d), xo = split->xs[i - 1];
            double x = position(box0, d);
            double xl, xu;
            if (xs < xo) { xl = xs; xu = xo; } else { xl = xo; xu = xs; }
            if (xl <= x && x <= xu) {
                if (!lfilled && xl < x) {
                    lower = std::fmax((xl + xu) / 2, (xl + x) / 2);
                    lfilled = true;
                } else if (!ufilled && xu > x) {
                    upper = std::fmin((xl + xu) / 2, (x + xu) / 2);
                    ufilled = true;
                }
            }
        }
        box = p;
    }
    return {lower, upper};
}

// src/tree.jl:184-216 (boxbounds!) -- separate restatement of the all-dimensions walk
inline std::vector<std::pair<double, double>> boxbounds(const Box* box) {
    int n = ndims(box);
    std::vector<char> lfilled(n, 0), ufilled(n, 0);
    std::vector<std::pair<double, double>> bbs(n);
    for (int d = 0; d < n; ++d) bbs[d] = {box->world->lower[d], box->world->upper[d]};
    const Box* box0 = box;
    int nlfilled = 0, nufilled = 0;
    while (!isroot(box) && std::min(nlfilled, nufilled) < n) {
        const Box* p = box->parent;
        const Split* split = p->split;
        for (size_t k = 0; k < split->dims.size(); ++k) {
            int sd = split->dims[k];
            double xo = split->xs[k];
            if (sd > n) continue;  // fictive dimensions
            if (!lfilled[sd - 1] || !ufilled[sd - 1]) {
                double xs = position(split->self, sd);
                double x = position(box0, sd);
                double xl, xu;
                if (xs < xo) { xl = xs; xu = xo; } else { xl = xo; xu = xs; }
                if (xl <= x && x <= xu) {
                    if (!lfilled[sd - 1] && xl < x) {
                        bbs[sd - 1].first = std::fmax((xl + xu) / 2, (xl + x) / 2);
                        lfilled[sd - 1] = 1;
                        nlfilled += 1;
                    } else if (!ufilled[sd - 1] && xu > x) {
                        bbs[sd - 1].second = std::fmin((xl + xu) / 2, (x + xu) / 2);
                        ufilled[sd - 1] = 1;
                        nufilled += 1;
                    }
                }
            }
        }
        box = p;
    }
    return bbs;
}

inline Box* getroot(Box* box) { while (!isroot(box)) box = box->parent; return box; }      // tree.jl:292-297
inline Box* getleaf(Box* box) { while (!isleaf(box)) box = box->split->self; return box; } // tree.jl:304-309

// src/tree.jl:316-323
inline double meta(const Box* box) {
    while (isself(box)) box = box->parent;
    if (isroot(box)) return box->world->meta;
    if (!(box->childindex > 0)) throw JlError(ERR_ASSERT, "AssertionError: box.childindex > 0");
    return box->parent->split->metas[box->childindex - 1];
}
inline double value(const Box* box) { return meta(box); }  // CoordinateSplittingPTrees.jl:61-62 (value(z::Real)=z)

// src/tree.jl:325-326
inline Box* getchild(const Split* split, int childidx) { return childidx == 0 ? split->self : split->others[childidx - 1]; }

// src/types.jl:203-246 (inner split constructor with its validity checks)
inline std::vector<Box*> split_box_full(Box* parent, const std::vector<int>& splitdims, const std::vector<double>& xs,
                                        const std::vector<double>& metas) {
    int p = parent->p, L = (1 << p) - 1;
    if ((int)splitdims.size() != p || (int)xs.size() != p || (int)metas.size() != L)
        throw JlError(ERR_DIMMISMATCH, "MethodError: split tuple lengths do not match the degree");
    if (!isleaf(parent)) throw JlError(ERR_ASSERT, "AssertionError: isleaf(parent)");
    int n = ndims(parent);
    int maxdim = ((n + p - 1) / p) * p;  // ceil(Int, n/p)*p
    for (int k = 0; k < p; ++k) {
        int splitdim = splitdims[k];
        double x = xs[k];
        if (!(0 < splitdim && splitdim <= maxdim)) throw JlError(ERR_ERROR, "got split along dimension beyond max allowed");
        if (splitdim > n) continue;
        auto bb = boxbounds(parent, splitdim);
        if ((x < bb.first) | (x > bb.second)) {
            throw JlError(ERR_ERROR, "out-of-bounds evaluation position");
        } else if (x == bb.second && bb.second != parent->world->upper[splitdim - 1]) {
            throw JlError(ERR_ERROR, "cannot split at the upper edge");
        }
        if (!(x != position(parent, splitdim))) throw JlError(ERR_ERROR, "position is identical to the parent");
    }
    Box* self = new Box{parent->world, parent, 0, nullptr, p};
    std::vector<Box*> others(L);
    for (int i = 1; i <= L; ++i) others[i - 1] = new Box{parent->world, parent, (uint8_t)i, nullptr, p};
    parent->split = new Split{splitdims, xs, self, others, metas};
    return others;
}

// src/types.jl:270-298 (outer constructors; k < p pads with fictive dimensions)
inline std::vector<Box*> split_box(Box* parent, const std::vector<int>& splitdims, const std::vector<double>& xs,
                                   const std::vector<double>& metas) {
    int p = parent->p, L = (1 << p) - 1;
    int k = (int)splitdims.size();
    if ((int)xs.size() != k) throw JlError(ERR_DIMMISMATCH, "MethodError: dims/xs length mismatch");
    if (k == p) return split_box_full(parent, splitdims, xs, metas);
    if (!(0 < k && k <= p)) throw JlError(ERR_DIMMISMATCH, "got k dimensions, max allowed is p");
    int nmeta = (1 << k) - 1;
    if ((int)metas.size() != nmeta) throw JlError(ERR_ERROR, "wrong number of metadatas for the split dimensions");
    int n = ndims(parent);
    double parentmeta = meta(parent);
    std::vector<int> dpad(p);
    std::vector<double> xpad(p), mpad(L);
    for (int d = 1; d <= p; ++d) {
        dpad[d - 1] = d <= k ? splitdims[d - 1] : n + d - k;
        xpad[d - 1] = d <= k ? xs[d - 1] : 0.0;
    }
    for (int d = 1; d <= L; ++d) {
        int i = d & nmeta;
        mpad[d - 1] = i == 0 ? parentmeta : metas[i - 1];
    }
    return split_box_full(parent, dpad, xpad, mpad);
}

// src/tree.jl:351-388
inline Box* find_leaf_at(Box* root, const double* x, int lenx) {
    int n = ndims(root);
    if (lenx != n) throw JlError(ERR_DIMMISMATCH, "tree has n dimensions, got a different length");
    for (int i = 1; i <= n; ++i) {
        auto bb = boxbounds(root, i);
        if (!((bb.first <= x[i - 1]) &&
              (x[i - 1] < bb.second || (x[i - 1] == bb.second && bb.second == root->world->upper[i - 1]))))
            throw JlError(ERR_ERROR, "x not within bounds along a dimension");
    }
    if (isleaf(root)) return root;
    while (!isleaf(root)) {
        const Split* split = root->split;
        int childidx = 0;
        for (int j = 1; j <= degree(root); ++j) {
            int sd = split->dims[j - 1];
            double xother = split->xs[j - 1];
            if (sd > n) continue;
            double xself = position(root, sd);
            double xsd = x[sd - 1];
            childidx |= (int)((xsd >= (xself + xother) / 2) ^ (xother < xself)) << (j - 1);
        }
        root = getchild(split, childidx);
    }
    return root;
}

// ---- preorder iteration, src/tree.jl:393-415, with VisitorState (types.jl:373-376)
struct VisitorState { Box* box; int childindex; };
struct BoxIterItem { bool ok; Box* item; VisitorState state; };

inline BoxIterItem box_iterate(Box* root) {
    if (isleaf(root)) return {true, root, {root, maxchildren(root)}};
    return {true, root, {root, 0}};
}
inline BoxIterItem box_iterate(Box* root, VisitorState state) {
    Box* box = state.box;
    int childindex = state.childindex;
    if (childindex < maxchildren(box) && !isleaf(box)) {
        Box* item = getchild(box->split, childindex);  // descend
        return {true, item, {item, isleaf(item) ? maxchildren(item) : 0}};
    }
    while (box != root) {  // ascend
        childindex = box->childindex + 1;
        box = box->parent;
        if (childindex < maxchildren(box)) {
            Box* item = getchild(box->split, childindex);
            return {true, item, {item, isleaf(item) ? maxchildren(item) : 0}};
        }
    }
    return {false, nullptr, {nullptr, 0}};
}
inline std::vector<Box*> collect(Box* root) {
    std::vector<Box*> out;
    for (auto r = box_iterate(root); r.ok; r = box_iterate(root, r.state)) out.push_back(r.item);
    return out;
}

// src/tree.jl:430-442
inline bool isfake(const Box* box) {
    int n = ndims(box);
    const Box* p = box->parent;
    if (isleaf(p)) return false;
    const Split* split = p->split;
    bool tf = false;
    unsigned cidx = box->childindex;
    for (int d : split->dims) {
        tf |= ((cidx & 1u) == 1u) & (d > n);
        cidx >>= 1;
    }
    return tf;
}
// src/tree.jl:423
inline std::vector<Box*> leaves(Box* root) {
    std::vector<Box*> out;
    for (auto r = box_iterate(root); r.ok; r = box_iterate(root, r.state))
        if (isleaf(r.item) && !isfake(r.item)) out.push_back(r.item);
    return out;
}
// Not in the reference: number every node / real leaf in the order of collect(root) / leaves(root).
inline void index_tree(Box* root) {
    int k = 0, l = 0;
    for (auto r = box_iterate(root); r.ok; r = box_iterate(root, r.state)) {
        r.item->preorder_id = k++;
        r.item->leaf_rank = (isleaf(r.item) && !isfake(r.item)) ? l++ : -1;
    }
}

// ---- Iterators.filter(isnonleaf, branchhead) over the preorder iterator (used at tree.jl:474,487,493,517,537)
struct BranchIter { Box* head; };
inline BoxIterItem filt_iterate(BranchIter it) {
    auto y = box_iterate(it.head);
    while (y.ok) {
        if (!isleaf(y.item)) return y;
        y = box_iterate(it.head, y.state);
    }
    return {false, nullptr, {nullptr, 0}};
}
inline BoxIterItem filt_iterate(BranchIter it, VisitorState st) {
    auto y = box_iterate(it.head, st);
    while (y.ok) {
        if (!isleaf(y.item)) return y;
        y = box_iterate(it.head, y.state);
    }
    return {false, nullptr, {nullptr, 0}};
}

// ---- splits(box): SplitIterator / ClimbingState, src/tree.jl:461-540, types.jl:393-405
struct ClimbingState {
    Box* box;
    bool visited;
    uint8_t skipchildindex;
    int childindex;
    BranchIter branchiter;
    VisitorState branchstate;
};
struct SplitIterItem { bool ok; Split* split; ClimbingState state; };

inline bool nobranch(Box* box, int skipchildindex, int childindex) {  // tree.jl:463-467
    return childindex < maxchildren(box) &&
           (childindex == skipchildindex || isleaf(getchild(box->split, childindex)));
}

inline SplitIterItem splits_iterate(Box* base) {  // tree.jl:469-496
    Box* box = base;
    const SplitIterItem none{false, nullptr, {nullptr, false, 0, 0, {nullptr}, {nullptr, 0}}};
    if (!isleaf(box)) {
        Box* branchhead = box->split->self;
        return {true, box->split, {box, true, 0xff, 0, {branchhead}, {branchhead, -1}}};
    }
    if (isroot(box)) return none;
    int skipchildindex = box->childindex, childindex = 0;
    box = box->parent;
    while (nobranch(box, skipchildindex, childindex)) childindex += 1;
    if (childindex >= maxchildren(box)) {
        BranchIter bi{box};
        auto r = filt_iterate(bi);
        return {true, r.item->split, {box, true, (uint8_t)skipchildindex, childindex, bi, r.state}};
    }
    Box* branchhead = getchild(box->split, childindex);
    BranchIter bi{branchhead};
    auto r = filt_iterate(bi);
    return {true, r.item->split, {box, false, (uint8_t)skipchildindex, childindex, bi, r.state}};
}

inline SplitIterItem splits_iterate(Box* /*base*/, const ClimbingState& state) {  // tree.jl:500-540
    const SplitIterItem none{false, nullptr, {nullptr, false, 0, 0, {nullptr}, {nullptr, 0}}};
    BranchIter branchiter = state.branchiter;
    VisitorState branchstate = state.branchstate;
    auto ret = branchstate.childindex < 0 ? filt_iterate(branchiter) : filt_iterate(branchiter, branchstate);
    if (ret.ok) {
        ClimbingState ns = state;
        ns.branchstate = ret.state;
        return {true, ret.item->split, ns};
    }
    Box* box = state.box;
    int skipchildindex = state.skipchildindex, childindex = state.childindex + 1;
    while (nobranch(box, skipchildindex, childindex)) childindex += 1;
    if (childindex < maxchildren(box)) {
        Box* branchhead = getchild(box->split, childindex);
        BranchIter bi{branchhead};
        auto r = filt_iterate(bi);
        return {true, r.item->split, {box, state.visited, (uint8_t)skipchildindex, childindex, bi, r.state}};
    }
    if (!state.visited)
        return {true, box->split, {box, true, state.skipchildindex, childindex, branchiter, branchstate}};
    if (isroot(box)) return none;
    skipchildindex = box->childindex;
    box = box->parent;
    childindex = 0;
    while (nobranch(box, skipchildindex, childindex)) childindex += 1;
    if (childindex >= maxchildren(box))
        return {true, box->split, {box, true, state.skipchildindex, childindex, branchiter, branchstate}};
    Box* branchhead = getchild(box->split, childindex);
    BranchIter bi{branchhead};
    auto r = filt_iterate(bi);
    return {true, r.item->split, {box, false, (uint8_t)skipchildindex, childindex, bi, r.state}};
}

// `state === nothing` (tree.jl:498) is modelled by has_state=false.
struct SplitCursor {
    Box* base;
    bool has_state = false;
    ClimbingState state{};
    SplitIterItem next() {
        SplitIterItem r = has_state ? splits_iterate(base, state) : splits_iterate(base);
        if (r.ok) { has_state = true; state = r.state; }
        return r;
    }
};
// src/tree.jl:580-587.  Destructuring `nothing` when the iterator runs dry is a MethodError in Julia.
inline SplitCursor skipsplits(Box* base, int skip) {
    SplitCursor c{base};
    if (skip == 0) return c;
    for (int i = 0; i < skip; ++i) {
        auto r = c.next();
        if (!r.ok) throw JlError(ERR_ERROR, "MethodError: iterator exhausted in skipsplits");
    }
    return c;
}
inline std::vector<Split*> collect_splits(Box* box) {
    std::vector<Split*> out;
    SplitCursor c{box};
    for (auto r = c.next(); r.ok; r = c.next()) out.push_back(r.split);
    return out;
}

// ---- chain(box), src/tree.jl:547-561
inline std::vector<Split*> collect_chain(Box* base) {
    std::vector<Split*> out;
    if (ndims(base) == 0) return out;
    Box* box = base;
    int count = 0;
    while (true) {
        if (isleaf(box)) throw JlError(ERR_UNDEFREF, "UndefRefError: access to undefined reference (box.split)");
        Split* item = box->split;
        out.push_back(item);
        box = item->others.back();
        count += degree(base);
        if (count >= ndims(base)) break;
    }
    return out;
}

// ---- chaintop, src/polynomials.jl:22-98 (restated with its control flow intact, dead third loop included)
inline std::pair<Box*, bool> chaintop(Box* box) {
    const int p = box->p;
    auto process_split = [p](std::vector<char>& covered, int nremaining, const Split* split, bool& didsplit) -> int {
        int n = (int)covered.size();
        int nnovel = p, ntarget = p;
        for (int d : split->dims) {
            nnovel -= (d > n || covered[d - 1]) ? 1 : 0;
            ntarget -= (d > n) ? 1 : 0;
        }
        if (!(nnovel == ntarget || nnovel == nremaining)) { didsplit = false; return nremaining; }
        for (int d : split->dims) {
            if (!(d <= n)) continue;
            covered[d - 1] = 1;
        }
        nremaining -= nnovel;
        didsplit = true;
        return nremaining;
    };
    int n = ndims(box);
    std::vector<char> covered(n, 0);
    int nremaining = n;
    Box *bottom = box, *top = box;
    bool didsplit = false;
    if (isleaf(box->parent)) throw JlError(ERR_UNDEFREF, "UndefRefError: access to undefined reference (box.parent.split)");
    const Split* split = box->parent->split;
    while (nremaining > 0) {  // First, go down
        nremaining = process_split(covered, nremaining, split, didsplit);
        if (!didsplit) break;
        bottom = split->others.back();
        if (isleaf(bottom)) break;
        bool hasfictive = false;
        for (int d : split->dims)
            if (d > n) hasfictive = true;
        if (hasfictive) break;
        split = bottom->split;
    }
    top = box->parent;
    if (nremaining == 0) return {top, true};
    if (isroot(box)) return {box, false};
    split = top->parent->split;
    while (nremaining > 0) {  // Go up
        nremaining = process_split(covered, nremaining, split, didsplit);
        if (!(didsplit || nremaining > 0)) break;
        if (isroot(top)) return {top, false};
        top = top->parent;
        split = top->parent->split;
    }
    while (nremaining > 0) {  // (unreachable in practice: see SURVEY.md M7)
        split = bottom->parent->split;
        for (int d : split->dims) {
            if (d < 1 || d > n) throw JlError(ERR_BOUNDS, "BoundsError: covered[d]");
            covered[d - 1] = 0;
        }
        nremaining += p;
        split = top->parent->split;
        nremaining = process_split(covered, nremaining, split, didsplit);
        if (!(didsplit || nremaining > 0)) break;
        if (isroot(top)) return {top, false};
        bottom = bottom->parent;
        top = top->parent;
    }
    return {top, nremaining == 0};
}

// ---- SymmetricArray (types.jl:432-454): N-d dense storage, getindex/setindex! sort the index tuple.
struct SymArray {
    int n = 0, N = 0;
    std::vector<double> data;  // column-major n^N
    SymArray() {}
    SymArray(int n_, int N_, double fill) : n(n_), N(N_) {
        size_t sz = 1;
        for (int i = 0; i < N; ++i) sz *= (size_t)n;
        data.assign(sz, fill);
    }
    size_t lin(std::vector<int> I) const {  // 1-based indices, sorted ascending (tuplesort, types.jl:451-454)
        for (size_t a = 1; a < I.size(); ++a)
            for (size_t b = a; b > 0 && I[b - 1] > I[b]; --b) std::swap(I[b - 1], I[b]);
        size_t off = 0, stride = 1;
        for (int k = 0; k < N; ++k) { off += (size_t)(I[k] - 1) * stride; stride *= (size_t)n; }
        return off;
    }
    double get(const std::vector<int>& I) const { return data[lin(I)]; }
    void set(const std::vector<int>& I, double v) { data[lin(I)] = v; }
    double get2(int i, int j) const { return i > j ? data[(size_t)(i - 1) * n + (j - 1)] : data[(size_t)(j - 1) * n + (i - 1)]; }
    void set2(int i, int j, double v) { if (i > j) std::swap(i, j); data[(size_t)(j - 1) * n + (i - 1)] = v; }
};

// ---- coefficients_p!, src/polynomials.jl:182-230 (dense SymmetricArray output, allocate :232-233, nnz :235-242)
struct CoefStats { long visited = 0, used = 0; };
inline SymArray coefficients_p(const std::function<double(const Box*)>& f, Box* box, int skip, CoefStats* stats = nullptr) {
    const int p = box->p;
    auto bitsum = [](unsigned i) {
        int s = (int)(i & 1u);
        while (i != 0) { i >>= 1; s += (int)(i & 1u); }
        return s;
    };
    auto bitsign = [&](unsigned i) { return (bitsum(i) & 1) ? -1 : 1; };
    int n = ndims(box);
    SymArray Cp(n, p, std::numeric_limits<double>::quiet_NaN());
    double nrem = n;  // nnz_offdiag_sym: n*(n-1)/2*(n-2)/3*...
    for (int i = 1; i <= p - 1; ++i) nrem *= (double)(n - i) / (double)(i + 1);
    long nremaining = std::lround(nrem);
    SplitCursor cur = skipsplits(box, skip);
    auto result = cur.next();
    while (result.ok) {
        Split* split = result.split;
        result = cur.next();
        if (stats) stats->visited++;
        const std::vector<int>& dims = split->dims;
        int mx = dims[0];
        for (int d : dims) mx = std::max(mx, d);
        if (mx > n) continue;
        if (std::isnan(Cp.get(dims))) {
            double denom = 1.0;
            for (size_t k = 0; k < dims.size(); ++k) denom *= split->xs[k] - position(split->self, dims[k]);
            int s = (p & 1) ? -1 : 1;
            double num = 0.0;
            for (int ichild = 0; ichild <= maxchildren(box) - 1; ++ichild) {
                Box* child = getchild(split, ichild);
                double fc = f(child);
                num += (double)(s * bitsign((unsigned)ichild)) * fc;
            }
            Cp.set(dims, num / denom);
            nremaining -= 1;
            if (stats) stats->used++;
        }
        if (nremaining <= 0) break;
    }
    return Cp;
}

// ---- coefficients_p! into a sparse / tridiagonal symmetric output (p = 2), src/polynomials.jl:182-230 with
// nnz_offdiag_sym :244-259 (stored entries above the diagonal; SymTridiagonal: its n-1 off-diagonals), fillnz! :261-271 and
// SymmetricArray getindex :440-443 (a structural zero reads 0.0, which is not NaN, so such pairs are skipped).
// `pattern` lists the stored above-diagonal positions (i < j, 1-based); out[k] is the coefficient of pattern[k] (NaN if never set).
inline std::vector<double> coefficients_p_pattern(const std::function<double(const Box*)>& f, Box* box, int skip,
                                                  const std::vector<std::pair<int, int>>& pattern, CoefStats* stats = nullptr) {
    if (box->p != 2) throw JlError(ERR_DIMMISMATCH, "DimensionMismatch: output should be 2-dimensional");
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    int n = ndims(box);
    std::vector<int> slot((size_t)n * n, -1);
    for (size_t k = 0; k < pattern.size(); ++k) {
        int i = pattern[k].first, j = pattern[k].second;
        if (!(1 <= i && i < j && j <= n)) throw JlError(ERR_ARGUMENT, "pattern entries must satisfy 1 <= i < j <= n");
        slot[(size_t)(i - 1) * n + (j - 1)] = (int)k;
    }
    std::vector<double> out(pattern.size(), NaN);
    long nremaining = (long)pattern.size();
    SplitCursor cur = skipsplits(box, skip);
    auto result = cur.next();
    while (result.ok) {
        Split* split = result.split;
        result = cur.next();
        if (stats) stats->visited++;
        int d1 = split->dims[0], d2 = split->dims[1];
        if (std::max(d1, d2) > n) continue;
        int lo = std::min(d1, d2), hi = std::max(d1, d2);
        int k = lo == hi ? -1 : slot[(size_t)(lo - 1) * n + (hi - 1)];
        if (k >= 0 && std::isnan(out[k])) {  // isnan(Cp[dims...])
            double denom = 1.0;
            for (size_t q = 0; q < 2; ++q) denom *= split->xs[q] - position(split->self, split->dims[q]);
            double num = 0.0;
            const int sgn[4] = {1, -1, -1, 1};
            for (int ichild = 0; ichild < 4; ++ichild) num += (double)sgn[ichild] * f(getchild(split, ichild));
            out[k] = num / denom;
            nremaining -= 1;
            if (stats) stats->used++;
        }
        if (nremaining <= 0) break;
    }
    return out;
}

// ---- chi / Qcoef helpers, src/cs2.jl:277-285, 302-325.  prec is n x n column-major bools: prec[i,j] -> prec[(j-1)*n+(i-1)]
struct Prec {
    int n = 0;
    std::vector<char> v;
    Prec() {}
    explicit Prec(int n_) : n(n_), v((size_t)n_ * n_, 0) {}
    bool at(int i, int j) const { return v[(size_t)(j - 1) * n + (i - 1)] != 0; }
    void setcol(int j, const std::vector<char>& col) { for (int i = 0; i < n; ++i) v[(size_t)(j - 1) * n + i] = col[i]; }
};
inline double chi(int j, int i, const double* b, const double* y, const Prec& prec) {  // cs2.jl:277-280
    if (i == j) return y[j - 1];
    return prec.at(i, j) ? b[j - 1] : 2 * y[j - 1] - b[j - 1];
}
inline double Qcoef_value(int i, int j, const double* x, const double* b, const double* y, const Prec& prec) {  // cs2.jl:302-312
    if (i == j) return (x[i - 1] - b[i - 1]) * (x[i - 1] - y[i - 1]) / 2;
    double bi = b[i - 1], bj = b[j - 1], yi = y[i - 1], yj = y[j - 1], xi = x[i - 1], xj = x[j - 1];
    bool pij = prec.at(i, j), pji = prec.at(j, i);
    if (((pij & !pji) & (xi == yi)) | ((pji & !pij) & (xj == yj))) return 0.0 * (xi * xj / 2);  // zero(xi*xj/2)
    double chiij = pji ? bi : 2 * yi - bi;
    double chiji = pij ? bj : 2 * yj - bj;
    return ((xi - bi) * (xj - chiji) + (xj - bj) * (xi - chiij)) / 4;
}
// cs2.jl:314-325.  NB the reference's line 317 parses as  A || (B && return coef):  when A holds, control falls
// through to the general expression (which is then analytically 0 but evaluated in floating point).
inline double Qcoef_lineq(int i, int k, double dxi, double xk, double bk, double yk, bool precik) {
    double coef = 0.0;
    if (i == k) {
        if (xk == bk && dxi == yk - bk) {
            // falls through
        } else if (xk == yk && dxi == bk - yk) {
            return coef;
        }
        return dxi / 2 + xk - (bk + yk) / 2;
    }
    if (precik) {
        if (xk == bk) return coef;
        return coef + (xk - bk);
    }
    return coef + (xk - yk);
}

// ---- fit_poly1, src/cs2.jl:162-205
struct Poly1 {
    double c = 0;
    std::vector<double> g, b, y;
    Prec prec;
    std::vector<int> firstmatch;
    Box* top = nullptr;
};
inline Poly1 fit_poly1(const std::function<double(const Box*)>& f, Box* box) {
    if (box->p != 2) throw JlError(ERR_ERROR, "MethodError: fit_poly1 is defined for Box{2} only");
    box = getleaf(box);
    int n = ndims(box);
    auto ct = chaintop(box);
    box = ct.first;
    if (!ct.second) throw JlError(ERR_ERROR, "no chain found");
    Poly1 r;
    r.top = box;
    r.b = position(box);
    r.y.assign(n, 0.0);  // `similar(b)`: uninitialised in Julia, every entry is overwritten below
    r.g.assign(n, 0.0);
    std::vector<char> wassplit(n, 0);
    r.firstmatch.assign(n, 0);
    r.prec = Prec(n);
    r.c = f(box);
    int npairs = (n + 1) / 2;
    for (int paircounter = 1; paircounter <= npairs; ++paircounter) {
        if (isleaf(box)) throw JlError(ERR_UNDEFREF, "UndefRefError: access to undefined reference (box.split)");
        Split* split = box->split;
        int d1 = split->dims[0], d2 = split->dims[1];
        if (!(d1 != d2)) throw JlError(ERR_ASSERT, "AssertionError: d1 != d2");
        if (d1 < 1 || d1 > n) throw JlError(ERR_BOUNDS, "BoundsError: wassplit[d1]");
        if (!(!wassplit[d1 - 1] && (d2 > n || !wassplit[d2 - 1])))
            throw JlError(ERR_ASSERT, "AssertionError: !(wassplit[d1]) && (d2 > n || !(wassplit[d2]))");
        double x1 = split->xs[0], x2 = split->xs[1];
        double f1 = f(split->others[0]), f2 = f(split->others[1]);
        (void)f(split->others[2]);  // f12 is read but unused by fit_poly1
        double f0 = f(split->self);
        r.y[d1 - 1] = x1; wassplit[d1 - 1] = 1; r.firstmatch[d1 - 1] = d2;
        if (d2 <= n) {
            r.y[d2 - 1] = x2; wassplit[d2 - 1] = 1; r.firstmatch[d2 - 1] = d1;
            r.prec.setcol(d2, wassplit);
        }
        r.prec.setcol(d1, wassplit);
        r.g[d1 - 1] = (f1 - f0) / (x1 - r.b[d1 - 1]);
        if (d2 <= n) r.g[d2 - 1] = (f2 - f0) / (x2 - r.b[d2 - 1]);
        box = split->others[2];
    }
    return r;
}

// ---- fit_poly2_diag!, src/cs2.jl:208-274
inline void fit_poly2_diag(const std::function<double(const Box*)>& f, SymArray& Q, Box* box, const Poly1& m, int skip,
                           long* nvisited = nullptr) {
    const std::vector<double>&g = m.g, &b = m.b, &y = m.y;
    const Prec& prec = m.prec;
    auto newinfo = [](double x, double bb, double yy) { return (x != bb) & (x != yy); };
    auto diag_equation = [&](int d, double xnew, double gd, const std::vector<double>& x, double& coef, double& consts) {
        int n = (int)x.size();
        double xd = x[d - 1];
        coef = Qcoef_lineq(d, d, xnew - xd, xd, b[d - 1], y[d - 1], prec.at(d, d));
        consts = -gd;
        for (int k = 1; k <= n; ++k) {
            if (k == d) continue;
            double xc = Qcoef_lineq(d, k, xnew - xd, x[k - 1], b[k - 1], y[k - 1], prec.at(d, k));
            consts -= Q.get2(d, k) * xc;
        }
    };
    auto process_split = [&](std::vector<char>& checked, Split* split, int dimidx, std::vector<double>& x) -> int {
        int n = (int)b.size();
        int d = split->dims[dimidx - 1];
        if (d > n || checked[d - 1]) return 0;
        double bd = b[d - 1], yd = y[d - 1];
        double xd = split->xs[dimidx - 1];
        if (newinfo(xd, bd, yd)) {
            double coef, consts;
            diag_equation(d, xd, g[d - 1], x, coef, consts);
            Box* child = split->others[dimidx - 1];
            Box* ref = split->self;
            Q.set2(d, d, (consts + (f(child) - f(ref)) / (xd - x[d - 1])) / coef);
            checked[d - 1] = 1;
            return 1;
        }
        if (newinfo(x[d - 1], bd, yd)) {  // try it the other way too
            std::swap(xd, x[d - 1]);
            int d2 = split->dims[3 - dimidx - 1];
            double xd2 = split->xs[3 - dimidx - 1];
            if (d2 <= n) std::swap(xd2, x[d2 - 1]);
            double coef, consts;
            diag_equation(d, xd, g[d - 1], x, coef, consts);
            Box* child = split->others[3 - dimidx - 1];
            Box* ref = split->others[2];
            Q.set2(d, d, (consts + (f(child) - f(ref)) / (xd - x[d - 1])) / coef);
            checked[d - 1] = 1;
            std::swap(xd, x[d - 1]);  // restore
            if (d2 <= n) std::swap(xd2, x[d2 - 1]);
            return 1;
        }
        return 0;
    };
    int n = ndims(box);
    std::vector<double> x = position(box);
    std::vector<char> filled(n, 0), checked(n, 0);
    int nremaining = n;
    long nsplits = 0;
    SplitCursor cur{box};
    for (auto r = cur.next(); r.ok; r = cur.next()) {
        if (nremaining <= 0) break;
        nsplits += 1;
        if (nvisited) (*nvisited)++;
        if (nsplits <= skip) continue;
        position_fill(x, filled, r.split->self);
        nremaining -= process_split(checked, r.split, 1, x);
        nremaining -= process_split(checked, r.split, 2, x);
    }
}

// ---- fit_quadratic_native / canonicalize / fit_quadratic, src/cs2.jl:89-93,115-121,132-140
struct NativeFit { Poly1 m; SymArray Q; };
inline NativeFit fit_quadratic_native(const std::function<double(const Box*)>& f, Box* box, int skip) {
    NativeFit r;
    r.m = fit_poly1(f, box);
    r.Q = coefficients_p(f, box, skip);
    fit_poly2_diag(f, r.Q, box, r.m, skip);
    return r;
}
inline std::vector<double> canonicalize(const std::vector<double>& g, const SymArray& Q, const std::vector<double>& b,
                                        const std::vector<double>& y, const Prec& prec) {
    std::vector<double> gc = g;
    int n = (int)g.size();
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
            double ch = chi(j, i, b.data(), y.data(), prec);
            gc[i - 1] += (b[j - 1] - ch) * Q.get2(i, j) / 2;
        }
    return gc;
}

// ---- modelvalue, native (cs2.jl:142-153) and canonical (cs2.jl:155-159)
inline double modelvalue_native(double c, const double* g, const SymArray& Q, const double* b, const double* y,
                                const Prec& prec, const double* x, int n) {
    double dot = 0;
    for (int i = 0; i < n; ++i) dot += g[i] * (x[i] - b[i]);
    double val = c + dot;
    for (int j = 1; j <= n; ++j)
        for (int i = j; i <= n; ++i) {
            double coef = Qcoef_value(i, j, x, b, y, prec);
            if (coef != 0) val += Q.get2(i, j) * coef * (double)(1 + (i != j));
        }
    return val;
}
// test/functions.jl:139-149
inline double modelvalue_chi(double c, const double* g, const SymArray& Q, const double* b, const double* y,
                             const Prec& prec, const double* x, int n) {
    double dot = 0;
    for (int i = 0; i < n; ++i) dot += g[i] * (x[i] - b[i]);
    double val = c + dot;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
            double coef = (x[i - 1] - b[i - 1]) * (x[j - 1] - chi(j, i, b, y, prec)) / 2;
            if (coef != 0) val += Q.get2(i, j) * coef;
        }
    return val;
}
// cs2.jl:155-159:  dx = x-b;  c + g'dx + (dx'Q dx)/2.  The reference's summation order is whatever LinearAlgebra's
// dot / generic matvec use (not pinned); this restatement forms (dx'Q) column by column and then dots with dx.
inline double modelvalue_canonical(double c, const double* g, const double* Qdata, const double* b, const double* x, int n) {
    // Qdata: SymmetricArray .data (column-major n x n); Q[i,j] reads data[min,max] (types.jl:440-443)
    std::vector<double> dx(n);  // `dx = x - b` allocates in the reference too
    for (int i = 0; i < n; ++i) dx[i] = x[i] - b[i];
    double gd = 0;
    for (int i = 0; i < n; ++i) gd += g[i] * dx[i];
    double quad = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += dx[i] * (i > j ? Qdata[(size_t)i * n + j] : Qdata[(size_t)j * n + i]);
        quad += s * dx[j];
    }
    return c + gd + quad / 2;
}
inline double modelvalue_canonical(double c, const double* g, const SymArray& Q, const double* b, const double* x, int n) {
    return modelvalue_canonical(c, g, Q.data.data(), b, x, n);
}

// ---- fit_quadratic_gdiag!, src/cs2.jl:332-384 ("next" row; restated for completeness of the checker)
inline void fit_quadratic_gdiag(const std::function<double(const Box*)>& f, SymArray& Q, Box* box, const std::vector<double>& b,
                                int skip, double& c, std::vector<double>& g) {
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    int n = ndims(box);
    g.assign(n, NaN);
    std::vector<double> qcoefs(2 * (size_t)n, NaN), rhs(2 * (size_t)n, NaN);  // 2 x n column-major
    std::vector<char> filled(n, 0);
    std::vector<double> x = position(box);
    int nremaining = 2 * n;
    c = f(box);
    long nsplits = 0;
    SplitCursor cur{box};
    for (auto r = cur.next(); r.ok; r = cur.next()) {
        if (nremaining == 0) break;
        nsplits += 1;
        if (nsplits <= skip) continue;
        Split* split = r.split;
        for (size_t i = 1; i <= split->dims.size(); ++i) {
            int d = split->dims[i - 1];
            if (!(d <= n)) continue;
            if (std::isnan(qcoefs[2 * (size_t)(d - 1) + 1])) {
                position_fill(x, filled, split->self);
                double xs = split->xs[i - 1];
                double xcoef = (x[d - 1] + xs) / 2 - b[d - 1];
                if (xcoef == qcoefs[2 * (size_t)(d - 1)]) continue;
                int idx = std::isnan(qcoefs[2 * (size_t)(d - 1)]) ? 1 : 2;
                double dx = xs - x[d - 1];
                double fd = (f(split->others[i - 1]) - f(split->self)) / dx;
                for (int j = 1; j <= n; ++j) {
                    if (j == d) continue;
                    fd -= Q.get2(d, j) * (x[j - 1] - b[j - 1]);
                }
                if (std::isfinite(xcoef) && std::isfinite(fd)) {
                    qcoefs[2 * (size_t)(d - 1) + (idx - 1)] = xcoef;
                    rhs[2 * (size_t)(d - 1) + (idx - 1)] = fd;
                    nremaining -= 1;
                }
            }
        }
    }
    for (int d = 1; d <= n; ++d) {
        if (!std::isnan(qcoefs[2 * (size_t)(d - 1) + 1])) {
            double a12 = qcoefs[2 * (size_t)(d - 1)], a22 = qcoefs[2 * (size_t)(d - 1) + 1];
            double r1 = rhs[2 * (size_t)(d - 1)], r2 = rhs[2 * (size_t)(d - 1) + 1];
            double det = a22 - a12;
            g[d - 1] = (a22 * r1 - a12 * r2) / det;
            Q.set2(d, d, (-r1 + r2) / det);
        }
    }
}

// ---- possemidef, src/posdef.jl:41-68.  Q is read through SymmetricArray getindex; the result is the full
// symmetric matrix `copy(Q)` with adjusted diagonal (column-major n x n).
inline std::vector<double> possemidef(const SymArray& Q) {
    const int n = Q.n;
    auto sgn = [](double q) { return q > 0 ? 1.0 : (q < 0 ? -1.0 : q); };                    // Julia sign()
    auto jlmax = [](double a, double b) { return (a != a || b != b) ? a + b : (a > b ? a : b); };  // Julia max(): NaN-propagating
    std::vector<double> Qp((size_t)n * n);
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) Qp[(size_t)(j - 1) * n + (i - 1)] = Q.get2(i, j);
    for (int i = 1; i <= n; ++i) Qp[(size_t)(i - 1) * n + (i - 1)] = 0;
    for (int j = 1; j <= n; ++j) {
        double qjj = std::fabs(Q.get2(j, j));
        for (int i = j + 1; i <= n; ++i) {
            double q = Q.get2(i, j);
            if ((q == 0) | (q != q)) continue;
            double qii = std::fabs(Q.get2(i, i));
            double r = sgn(q) * std::sqrt(qjj / qii);
            if (r == 0 || !std::isfinite(r)) r = sgn(q);
            Qp[(size_t)(j - 1) * n + (j - 1)] += q * r;
            Qp[(size_t)(i - 1) * n + (i - 1)] += q / r;
        }
    }
    for (int i = 1; i <= n; ++i) {
        double& d = Qp[(size_t)(i - 1) * n + (i - 1)];
        d = jlmax(d, std::fabs(Q.get2(i, i)));
    }
    return Qp;
}

// ---- addpoint! (4-argument form), src/CoordinateSplittingPTrees.jl:94-142
inline Box* addpoint(Box* box, const double* x, int lenx, const std::vector<std::vector<int>>& dimlists, const Objective& metagen) {
    int n = ndims(box);
    std::vector<char> covered(n, 0);
    for (const auto& dimlist : dimlists)
        for (int d : dimlist) {
            if (d < 1 || d > n) throw JlError(ERR_BOUNDS, "BoundsError: covered[d]");
            if (covered[d - 1]) throw JlError(ERR_ERROR, "dimension was duplicated");
            covered[d - 1] = 1;
        }
    for (int d = 0; d < n; ++d)
        if (!covered[d]) throw JlError(ERR_ERROR, "all dimensions must be covered");
    if (!isleaf(box)) box = find_leaf_at(box, x, lenx);
    if (lenx != n) throw JlError(ERR_BOUNDS, "BoundsError: x[d]");
    std::vector<double> b = position(box);
    std::vector<double> y = b;
    std::vector<double> metas;
    for (const auto& dimlist : dimlists) {
        int k = (int)dimlist.size();
        metas.clear();
        for (unsigned i = 1; i <= (1u << k) - 1; ++i) {
            std::vector<double> y0(k);
            for (int j = 0; j < k; ++j)
                if ((i >> j) & 1u) { y0[j] = y[dimlist[j] - 1]; y[dimlist[j] - 1] = x[dimlist[j] - 1]; }
            metas.push_back(metagen(y.data(), n));
            for (int j = 0; j < k; ++j)
                if ((i >> j) & 1u) y[dimlist[j] - 1] = y0[j];
        }
        std::vector<double> xs(k);
        for (int j = 0; j < k; ++j) xs[j] = x[dimlist[j] - 1];
        std::vector<Box*> others = split_box(box, dimlist, xs, metas);
        unsigned cindex = 0;
        for (int d : dimlist) cindex = (cindex << 1) | (unsigned)(d <= n);
        box = others[cindex - 1];
        for (int j = 0; j < k; ++j) y[dimlist[j] - 1] = x[dimlist[j] - 1];
    }
    return box;
}

inline void free_tree(Box* root) {
    std::vector<Box*> all = collect(root);
    for (Box* b : all) { delete b->split; delete b; }
}

}  // namespace cspo
