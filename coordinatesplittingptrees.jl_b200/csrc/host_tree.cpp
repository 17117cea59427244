// host_tree.cpp -- libcsp_host.so: the host side of the drop-in, standing in for the Julia object graph.
//
// In the reference the tree is grown on the host by `addpoint!` (src/CoordinateSplittingPTrees.jl:94-142) into a graph
// of mutable `Box` structs (src/types.jl:187-246); growth stays on the host in this design too.  Julia is not installed
// in this image, so the host side that a Julia wrapper would provide (World / Box / addpoint! / flatten-to-SoA) is
// written here in C++, behind a small C ABI used by the Python mirror in ../tree.py.  It is NOT a port of the reference's
// data structure: boxes live in flat growable arrays, every box stores its position and value (so no walks to the
// root), and box bounds are derived top-down.  The flattener emits exactly the csp_tree_desc arrays of
// include/csp_b200.h in preorder (the order of `collect(root)`, src/tree.jl:393-415).
//
// Error codes mirror the Julia exception classes the reference throws for the same misuse:
//   -1 ErrorException, -2 DimensionMismatch, -3 AssertionError, -5 ArgumentError, -6 BoundsError
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }
enum { E_ERROR = -1, E_DIMMISMATCH = -2, E_ASSERT = -3, E_ARGUMENT = -5, E_BOUNDS = -6 };

typedef double (*objective_fn)(const double* y, int n, void* ctx);

struct HostTree {
    int n = 0, p = 0, C = 0;
    std::vector<double> lower, upper;
    // per box, in creation order (ids are stable under growth)
    std::vector<int> parent;        // root: itself
    std::vector<uint8_t> childindex;
    std::vector<int> first_child;   // -1: leaf; children are first_child .. first_child + C - 1
    std::vector<int> split_id;      // -1: leaf
    std::vector<double> val;        // value(box)
    std::vector<double> pos;        // n per box: position(box)
    // per split
    std::vector<int> sdims;         // p per split, 1-based, > n fictive
    std::vector<double> sxs;        // p per split
    std::vector<int> sbox;          // the box each split divides: with sdims / sxs / the children's values this is the split log
    int n_fake_leaves = 0;

    int nboxes() const { return (int)parent.size(); }
    bool isleaf(int b) const { return first_child[b] < 0; }
    int add_box(int par, int ci, double v, const double* x) {
        int id = nboxes();
        parent.push_back(par < 0 ? id : par);
        childindex.push_back((uint8_t)ci);
        first_child.push_back(-1);
        split_id.push_back(-1);
        val.push_back(v);
        pos.insert(pos.end(), x, x + n);
        return id;
    }
    bool isfake(int b) const {  // displaced along a fictive dimension (reference src/tree.jl:430-442)
        int par = parent[b];
        if (par == b) return false;
        int s = split_id[par];
        unsigned ci = childindex[b];
        for (int k = 0; k < p; ++k)
            if (((ci >> k) & 1u) && sdims[(size_t)s * p + k] > n) return true;
        return false;
    }
    // bounds of `box` along 1-based dimension d, derived top-down from the splits on the root path
    void boxbounds(int box, int d, double& lo, double& hi) const {
        lo = lower[d - 1]; hi = upper[d - 1];
        std::vector<int> path;
        for (int b = box; parent[b] != b; b = parent[b]) path.push_back(b);
        for (auto it = path.rbegin(); it != path.rend(); ++it) {
            int b = *it, par = parent[b], s = split_id[par];
            for (int k = 0; k < p; ++k) {
                if (sdims[(size_t)s * p + k] != d) continue;
                double xself = pos[(size_t)par * n + d - 1], xo = sxs[(size_t)s * p + k];
                double mid = (xself + xo) / 2;
                bool displaced = (childindex[b] >> k) & 1u;
                double mine = displaced ? xo : xself, other = displaced ? xself : xo;
                if (mine < other) hi = mid; else lo = mid;
            }
        }
    }
    // descent from `from` (reference semantics of src/tree.jl:362-386)
    int descend(int from, const double* x) const {
        int b = from;
        while (!isleaf(b)) {
            int s = split_id[b];
            int idx = 0;
            for (int k = 0; k < p; ++k) {
                int sd = sdims[(size_t)s * p + k];
                if (sd > n) continue;
                double xother = sxs[(size_t)s * p + k], xself = pos[(size_t)b * n + sd - 1];
                idx |= (int)((x[sd - 1] >= (xself + xother) / 2) ^ (xother < xself)) << k;
            }
            b = first_child[b] + idx;
        }
        return b;
    }
};

int do_find_leaf(const HostTree& t, int from, const double* x, int lenx, int* out) {
    if (lenx != t.n) return fail(E_DIMMISMATCH, "tree dimensionality and length(x) differ");
    for (int i = 1; i <= t.n; ++i) {
        double lo, hi;
        t.boxbounds(from, i, lo, hi);
        double xi = x[i - 1];
        if (!((lo <= xi) && (xi < hi || (xi == hi && hi == t.upper[i - 1])))) return fail(E_ERROR, "x is not within the box bounds");
    }
    *out = t.descend(from, x);
    return 0;
}

// Box(parent, dims, xs, metas) with the reference's validity rules (src/types.jl:203-246, padding :278-298)
int do_split(HostTree& t, int par, int k, const int* dims, const double* xs, int nmeta, const double* metas, int* others_out) {
    const int n = t.n, p = t.p, L = t.C - 1;
    if (!(0 < k && k <= p)) return fail(E_DIMMISMATCH, "number of split dimensions must be between 1 and p");
    if (nmeta != (1 << k) - 1) return fail(k == p ? E_DIMMISMATCH : E_ERROR, "wrong number of metadata values for the split dimensions");
    if (!t.isleaf(par)) return fail(E_ASSERT, "isleaf(parent)");
    std::vector<int> d(p);
    std::vector<double> x(p), m(L);
    for (int j = 1; j <= p; ++j) {
        d[j - 1] = j <= k ? dims[j - 1] : n + j - k;
        x[j - 1] = j <= k ? xs[j - 1] : 0.0;
    }
    for (int c = 1; c <= L; ++c) {
        int i = c & nmeta;
        m[c - 1] = i == 0 ? t.val[par] : metas[i - 1];
    }
    const int maxdim = (n + p - 1) / p * p;
    for (int j = 0; j < p; ++j) {
        if (!(0 < d[j] && d[j] <= maxdim)) return fail(E_ERROR, "split dimension beyond the maximum allowed");
        if (d[j] > n) continue;
        double lo, hi;
        t.boxbounds(par, d[j], lo, hi);
        if ((x[j] < lo) | (x[j] > hi)) return fail(E_ERROR, "out-of-bounds evaluation position");
        if (x[j] == hi && hi != t.upper[d[j] - 1]) return fail(E_ERROR, "cannot split at the upper edge of a box that is not at the world edge");
        if (!(x[j] != t.pos[(size_t)par * n + d[j] - 1])) return fail(E_ERROR, "position is identical to the parent");
    }
    if (t.isfake(par)) t.n_fake_leaves--;  // a fake leaf that gets split becomes an internal box (counted there from now on)
    const int s = (int)(t.sdims.size() / p);
    t.sdims.insert(t.sdims.end(), d.begin(), d.end());
    t.sxs.insert(t.sxs.end(), x.begin(), x.end());
    t.sbox.push_back(par);
    std::vector<double> cp(t.pos.begin() + (size_t)par * n, t.pos.begin() + (size_t)(par + 1) * n);
    const int first = t.nboxes();
    for (int c = 0; c <= L; ++c) {
        std::vector<double> xc = cp;
        bool fake = false;
        for (int j = 0; j < p; ++j)
            if ((c >> j) & 1) {
                if (d[j] <= n) xc[d[j] - 1] = x[j];
                else fake = true;
            }
        int id = t.add_box(par, c, c == 0 ? t.val[par] : m[c - 1], xc.data());
        if (fake) t.n_fake_leaves++;
        if (c > 0 && others_out) others_out[c - 1] = id;
    }
    t.first_child[par] = first;
    t.split_id[par] = s;
    return 0;
}

// addpoint!(box, x, dimlists, metagen) (reference src/CoordinateSplittingPTrees.jl:94-142)
int do_addpoint(HostTree& t, int box, const double* x, int lenx, int ndl, const int* dl_flat, const int* dl_len, objective_fn f,
                void* ctx, int* out) {
    const int n = t.n;
    std::vector<char> covered(n, 0);
    for (int i = 0, o = 0; i < ndl; ++i)
        for (int j = 0; j < dl_len[i]; ++j, ++o) {
            int d = dl_flat[o];
            if (d < 1 || d > n) return fail(E_BOUNDS, "dimension in dimlists out of range");
            if (covered[d - 1]) return fail(E_ERROR, "a dimension was duplicated in dimlists");
            covered[d - 1] = 1;
        }
    for (int d = 0; d < n; ++d)
        if (!covered[d]) return fail(E_ERROR, "all dimensions must be covered by dimlists");
    if (!t.isleaf(box)) {
        int rc = do_find_leaf(t, box, x, lenx, &box);
        if (rc) return rc;
    }
    if (lenx != n) return fail(E_BOUNDS, "length(x) differs from the tree dimensionality");
    std::vector<double> y(t.pos.begin() + (size_t)box * n, t.pos.begin() + (size_t)(box + 1) * n);
    std::vector<double> metas, xs;
    std::vector<int> others(t.C - 1);
    for (int i = 0, o = 0; i < ndl; o += dl_len[i], ++i) {
        const int k = dl_len[i];
        const int* dl = dl_flat + o;
        if (k < 1 || k > t.p || k > 8) return fail(E_DIMMISMATCH, "number of split dimensions must be between 1 and p");
        metas.clear();
        for (unsigned c = 1; c <= (1u << k) - 1; ++c) {  // corner positions in bit-pattern order
            double save[8];
            for (int j = 0; j < k; ++j)
                if ((c >> j) & 1u) { save[j] = y[dl[j] - 1]; y[dl[j] - 1] = x[dl[j] - 1]; }
            metas.push_back(f(y.data(), n, ctx));
            for (int j = 0; j < k; ++j)
                if ((c >> j) & 1u) y[dl[j] - 1] = save[j];
        }
        xs.resize(k);
        for (int j = 0; j < k; ++j) xs[j] = x[dl[j] - 1];
        int rc = do_split(t, box, k, dl, xs.data(), (int)metas.size(), metas.data(), others.data());
        if (rc) return rc;
        box = others[(1 << k) - 2];  // the child displaced along every real dimension of the list
        for (int j = 0; j < k; ++j) y[dl[j] - 1] = x[dl[j] - 1];
    }
    *out = box;
    return 0;
}

}  // namespace

extern "C" {

const char* csph_last_error() { return g_err.c_str(); }

int csph_tree_new(int n, int p, const double* lower, const double* upper, const double* position, double meta, void** out) {
    if (n < 1) return fail(E_DIMMISMATCH, "n must be >= 1");
    if (p < 1 || p > 8) return fail(E_ERROR, "degree must be between 1 and 8");
    for (int i = 0; i < n; ++i)
        if (!(lower[i] <= position[i] && position[i] <= upper[i])) return fail(E_ARGUMENT, "position is not within bounds");
    HostTree* t = new HostTree();
    t->n = n; t->p = p; t->C = 1 << p;
    t->lower.assign(lower, lower + n);
    t->upper.assign(upper, upper + n);
    t->add_box(-1, 0, meta, position);
    *out = t;
    return 0;
}
void csph_tree_free(void* t) { delete (HostTree*)t; }

int csph_split(void* tp, int parent, int k, const int* dims, const double* xs, int nmeta, const double* metas, int* others_out) {
    HostTree& t = *(HostTree*)tp;
    if (parent < 0 || parent >= t.nboxes()) return fail(E_BOUNDS, "box id out of range");
    return do_split(t, parent, k, dims, xs, nmeta, metas, others_out);
}
int csph_addpoint(void* tp, int box, const double* x, int lenx, int ndl, const int* dl_flat, const int* dl_len, objective_fn f,
                  void* ctx, int* out) {
    HostTree& t = *(HostTree*)tp;
    if (box < 0 || box >= t.nboxes()) return fail(E_BOUNDS, "box id out of range");
    return do_addpoint(t, box, x, lenx, ndl, dl_flat, dl_len, f, ctx, out);
}
// one addpoint! from the root per row of X (npts x n), dimlists taken cyclically from `ncycle` lists-of-lists
int csph_addpoints_cycle(void* tp, const double* X, long npts, int n, int ncycle, const int* ndl_per, const int* dl_flat,
                         const int* dl_len, objective_fn f, void* ctx) {
    HostTree& t = *(HostTree*)tp;
    if (n != t.n) return fail(E_DIMMISMATCH, "X has the wrong number of columns");
    std::vector<int> flat_off(ncycle), len_off(ncycle);
    for (int c = 0, fo = 0, lo = 0; c < ncycle; ++c) {
        flat_off[c] = fo; len_off[c] = lo;
        for (int i = 0; i < ndl_per[c]; ++i) fo += dl_len[lo + i];
        lo += ndl_per[c];
    }
    for (long i = 0; i < npts; ++i) {
        int c = (int)(i % ncycle), out;
        int rc = do_addpoint(t, 0, X + (size_t)i * n, n, ndl_per[c], dl_flat + flat_off[c], dl_len + len_off[c], f, ctx, &out);
        if (rc) return rc;
    }
    return 0;
}

int csph_find_leaf_host(void* tp, int from, const double* x, int lenx, int* out) {
    // growth-time helper only (addpoint! needs the target leaf); batched queries go through libcsp_b200.so
    HostTree& t = *(HostTree*)tp;
    if (from < 0 || from >= t.nboxes()) return fail(E_BOUNDS, "box id out of range");
    return do_find_leaf(t, from, x, lenx, out);
}

int csph_counts(void* tp, int* n, int* p, int* n_boxes, int* n_internal, int* n_leaves) {
    HostTree& t = *(HostTree*)tp;
    int ni = (int)(t.sdims.size() / t.p);
    *n = t.n; *p = t.p; *n_boxes = t.nboxes(); *n_internal = ni;
    *n_leaves = t.nboxes() - ni - t.n_fake_leaves;
    return 0;
}
int csph_box_info(void* tp, int box, int* parent, int* childindex, int* first_child, int* isfake, double* val, double* pos,
                  int* dims, double* xs) {
    HostTree& t = *(HostTree*)tp;
    if (box < 0 || box >= t.nboxes()) return fail(E_BOUNDS, "box id out of range");
    *parent = t.parent[box]; *childindex = t.childindex[box]; *first_child = t.first_child[box];
    *isfake = t.isfake(box) ? 1 : 0;
    *val = t.val[box];
    if (pos) std::copy(t.pos.begin() + (size_t)box * t.n, t.pos.begin() + (size_t)(box + 1) * t.n, pos);
    if (!t.isleaf(box)) {
        int s = t.split_id[box];
        for (int k = 0; k < t.p; ++k) {
            if (dims) dims[k] = t.sdims[(size_t)s * t.p + k];
            if (xs) xs[k] = t.sxs[(size_t)s * t.p + k];
        }
    }
    return 0;
}
int csph_boxbounds(void* tp, int box, int d, double* out2) {
    HostTree& t = *(HostTree*)tp;
    if (box < 0 || box >= t.nboxes()) return fail(E_BOUNDS, "box id out of range");
    if (d < 1 || d > t.n) return fail(E_BOUNDS, "dimension out of range");
    t.boxbounds(box, d, out2[0], out2[1]);
    return 0;
}

// The split log (creation order), entries [s0, s1): what csp_tree_append_splits consumes.  The children of the s-th split are
// the boxes 1 + s*2^p + (0 .. 2^p-1); child_val[(s-s0)*2^p + c] = value(child c).
int csph_split_count(void* tp, int* nsplits) {
    HostTree& t = *(HostTree*)tp;
    *nsplits = (int)t.sbox.size();
    return 0;
}
int csph_split_log(void* tp, int s0, int s1, int* split_box, int* dims, double* xs, double* child_val) {
    HostTree& t = *(HostTree*)tp;
    const int p = t.p, C = t.C;
    if (s0 < 0 || s1 < s0 || s1 > (int)t.sbox.size()) return fail(E_BOUNDS, "split range out of bounds");
    for (int s = s0; s < s1; ++s) {
        const int k = s - s0, par = t.sbox[s], first = t.first_child[par];
        if (first != 1 + s * C) return fail(E_ERROR, "internal error: children of a split are not at 1 + s*2^p");
        split_box[k] = par;
        for (int j = 0; j < p; ++j) { dims[(size_t)k * p + j] = t.sdims[(size_t)s * p + j]; xs[(size_t)k * p + j] = t.sxs[(size_t)s * p + j]; }
        for (int c = 0; c < C; ++c) child_val[(size_t)k * C + c] = t.val[first + c];
    }
    return 0;
}

// Preorder flatten into the csp_tree_desc arrays (caller allocates from csph_counts).  Also returns the two id maps:
// pre_of_box[box id] = node id, box_of_pre[node id] = box id.
int csph_flatten(void* tp, double* lower, double* upper, int* parent, uint8_t* childindex, int* internal_id, int* leaf_id,
                 double* val, int* inode, int* dims, double* xs, int* child, double* pos, int* pre_of_box, int* box_of_pre) {
    HostTree& t = *(HostTree*)tp;
    const int n = t.n, p = t.p, C = t.C, nb = t.nboxes();
    std::copy(t.lower.begin(), t.lower.end(), lower);
    std::copy(t.upper.begin(), t.upper.end(), upper);
    std::vector<int> stack;
    stack.reserve(256);
    stack.push_back(0);
    int v = 0;
    while (!stack.empty()) {  // pass 1: preorder numbering
        int b = stack.back();
        stack.pop_back();
        pre_of_box[b] = v;
        box_of_pre[v] = b;
        ++v;
        if (!t.isleaf(b))
            for (int k = C - 1; k >= 0; --k) stack.push_back(t.first_child[b] + k);
    }
    if (v != nb) return fail(E_ERROR, "internal error: preorder walk did not reach every box");
    int I = 0, L = 0;
    for (int u = 0; u < nb; ++u) {
        int b = box_of_pre[u];
        parent[u] = pre_of_box[t.parent[b]];
        childindex[u] = t.childindex[b];
        val[u] = t.val[b];
        if (t.isleaf(b)) {
            internal_id[u] = -1;
            leaf_id[u] = t.isfake(b) ? -1 : L++;
            continue;
        }
        internal_id[u] = I;
        leaf_id[u] = -1;
        inode[I] = u;
        int s = t.split_id[b];
        for (int k = 0; k < p; ++k) { dims[(size_t)I * p + k] = t.sdims[(size_t)s * p + k]; xs[(size_t)I * p + k] = t.sxs[(size_t)s * p + k]; }
        for (int k = 0; k < C; ++k) child[(size_t)I * C + k] = pre_of_box[t.first_child[b] + k];
        std::copy(t.pos.begin() + (size_t)b * n, t.pos.begin() + (size_t)(b + 1) * n, pos + (size_t)I * n);
        ++I;
    }
    return 0;
}

}  // extern "C"
