// capi.cpp -- C entry points over csp_oracle.hpp for ctypes.  TEST INFRASTRUCTURE ONLY (see csp_oracle.hpp).
// Every function returns 0 on success or -ErrKind (the Julia exception class the reference would throw);
// cspo_last_error() gives the message.  Box* / World* cross the boundary as opaque pointers.
#include "csp_oracle.hpp"

#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace cspo;

static thread_local std::string g_err;
static int fail(const JlError& e) { g_err = e.what(); return -e.kind; }
#define CSPO_TRY try {
#define CSPO_CATCH                                                         \
    }                                                                      \
    catch (const JlError& e) { return fail(e); }                           \
    catch (const std::exception& e) { g_err = e.what(); return -ERR_ERROR; } \
    return 0;

typedef double (*cspo_objective)(const double* y, int n, void* ctx);

static const std::function<double(const Box*)> kValue = [](const Box* b) { return value(b); };

extern "C" {

const char* cspo_last_error() { return g_err.c_str(); }
int cspo_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int cspo_world_new(int n, const double* lo, const double* up, const double* pos, double meta, void** out) {
    CSPO_TRY
    *out = make_world(std::vector<double>(lo, lo + n), std::vector<double>(up, up + n), std::vector<double>(pos, pos + n), meta);
    CSPO_CATCH
}
void cspo_world_free(void* w) { delete (World*)w; }
int cspo_root_new(void* world, int p, void** out) {
    CSPO_TRY
    *out = make_root((World*)world, p);
    CSPO_CATCH
}
void cspo_tree_free(void* root) { free_tree((Box*)root); }

int cspo_split(void* parent, int k, const int* dims, const double* xs, int nmeta, const double* metas, void** others_out) {
    CSPO_TRY
    auto others = split_box((Box*)parent, std::vector<int>(dims, dims + k), std::vector<double>(xs, xs + k),
                            std::vector<double>(metas, metas + nmeta));
    for (size_t i = 0; i < others.size(); ++i) others_out[i] = others[i];
    CSPO_CATCH
}

int cspo_addpoint(void* box, const double* x, int lenx, int ndl, const int* dl_flat, const int* dl_len, cspo_objective f,
                  void* ctx, void** out) {
    CSPO_TRY
    std::vector<std::vector<int>> dimlists;
    int o = 0;
    for (int i = 0; i < ndl; ++i) {
        dimlists.emplace_back(dl_flat + o, dl_flat + o + dl_len[i]);
        o += dl_len[i];
    }
    *out = addpoint((Box*)box, x, lenx, dimlists, [f, ctx](const double* y, int n) { return f(y, n, ctx); });
    CSPO_CATCH
}

// Grow with a fixed cycle of dimlists (ncycle lists-of-lists), one addpoint! per row of X (column-major n x npts).
int cspo_addpoints_cycle(void* root, const double* X, long npts, int n, int ncycle, const int* ndl_per, const int* dl_flat,
                         const int* dl_len, cspo_objective f, void* ctx) {
    CSPO_TRY
    std::vector<std::vector<std::vector<int>>> cyc(ncycle);
    int o = 0, li = 0;
    for (int c = 0; c < ncycle; ++c)
        for (int i = 0; i < ndl_per[c]; ++i) {
            cyc[c].emplace_back(dl_flat + o, dl_flat + o + dl_len[li]);
            o += dl_len[li++];
        }
    Objective obj = [f, ctx](const double* y, int nn) { return f(y, nn, ctx); };
    for (long t = 0; t < npts; ++t) addpoint((Box*)root, X + (size_t)t * n, n, cyc[t % ncycle], obj);
    CSPO_CATCH
}

int cspo_isleaf(void* b) { return isleaf((Box*)b); }
int cspo_isroot(void* b) { return isroot((Box*)b); }
int cspo_isfake(void* b) { return isfake((Box*)b); }
int cspo_ndims(void* b) { return ndims((Box*)b); }
int cspo_degree(void* b) { return degree((Box*)b); }
void* cspo_parent(void* b) { return ((Box*)b)->parent; }
int cspo_childindex(void* b) { return ((Box*)b)->childindex; }
void* cspo_getroot(void* b) { return getroot((Box*)b); }
void* cspo_getleaf(void* b) { return getleaf((Box*)b); }
int cspo_getchild(void* b, int idx, void** out) {
    CSPO_TRY
    Box* box = (Box*)b;
    if (isleaf(box)) throw JlError(ERR_UNDEFREF, "UndefRefError: box.split");
    if (idx < 0 || idx >= maxchildren(box)) throw JlError(ERR_BOUNDS, "BoundsError: child index");
    *out = getchild(box->split, idx);
    CSPO_CATCH
}
int cspo_split_info(void* b, int* dims, double* xs) {
    CSPO_TRY
    Box* box = (Box*)b;
    if (isleaf(box)) throw JlError(ERR_UNDEFREF, "UndefRefError: box.split");
    for (int i = 0; i < box->p; ++i) { dims[i] = box->split->dims[i]; xs[i] = box->split->xs[i]; }
    CSPO_CATCH
}
int cspo_position_d(void* b, int d, double* out) {
    CSPO_TRY
    *out = position((Box*)b, d);
    CSPO_CATCH
}
int cspo_position(void* b, double* out) {
    CSPO_TRY
    auto x = position((Box*)b);
    std::copy(x.begin(), x.end(), out);
    CSPO_CATCH
}
int cspo_boxbounds_d(void* b, int d, double* out2) {
    CSPO_TRY
    auto bb = boxbounds((Box*)b, d);
    out2[0] = bb.first; out2[1] = bb.second;
    CSPO_CATCH
}
int cspo_boxbounds(void* b, double* out2n) {
    CSPO_TRY
    auto bbs = boxbounds((Box*)b);
    for (size_t i = 0; i < bbs.size(); ++i) { out2n[2 * i] = bbs[i].first; out2n[2 * i + 1] = bbs[i].second; }
    CSPO_CATCH
}
int cspo_meta(void* b, double* out) {
    CSPO_TRY
    *out = meta((Box*)b);
    CSPO_CATCH
}
int cspo_find_leaf_at(void* root, const double* x, int lenx, void** out) {
    CSPO_TRY
    *out = find_leaf_at((Box*)root, x, lenx);
    CSPO_CATCH
}
long cspo_collect(void* root, void** out, long cap) {
    auto v = collect((Box*)root);
    for (long i = 0; i < (long)v.size() && i < cap; ++i) out[i] = v[i];
    return (long)v.size();
}
long cspo_leaves(void* root, void** out, long cap) {
    auto v = leaves((Box*)root);
    for (long i = 0; i < (long)v.size() && i < cap; ++i) out[i] = v[i];
    return (long)v.size();
}
// splits(box): returns the boxes that own each yielded split (split.self.parent), in iteration order.
long cspo_splits(void* box, void** out, long cap) {
    auto v = collect_splits((Box*)box);
    for (long i = 0; i < (long)v.size() && i < cap; ++i) out[i] = v[i]->self->parent;
    return (long)v.size();
}
int cspo_chain(void* box, void** out, long cap, long* count) {
    CSPO_TRY
    auto v = collect_chain((Box*)box);
    for (long i = 0; i < (long)v.size() && i < cap; ++i) out[i] = v[i]->self->parent;
    *count = (long)v.size();
    CSPO_CATCH
}
int cspo_chaintop(void* box, void** top, int* success) {
    CSPO_TRY
    auto r = chaintop((Box*)box);
    *top = r.first;
    *success = r.second;
    CSPO_CATCH
}
void cspo_index_tree(void* root) { index_tree((Box*)root); }
int cspo_preorder_id(void* b) { return ((Box*)b)->preorder_id; }
int cspo_leaf_rank(void* b) { return ((Box*)b)->leaf_rank; }

int cspo_coefficients_p(void* box, int skip, double* out, long* visited, long* used) {
    CSPO_TRY
    CoefStats st;
    SymArray Cp = coefficients_p(kValue, (Box*)box, skip, &st);
    std::copy(Cp.data.begin(), Cp.data.end(), out);
    if (visited) *visited = st.visited;
    if (used) *used = st.used;
    CSPO_CATCH
}

static void copy_native(const NativeFit& r, double* c, double* g, double* Q, double* b, double* y, unsigned char* prec, int* firstmatch) {
    *c = r.m.c;
    std::copy(r.m.g.begin(), r.m.g.end(), g);
    std::copy(r.Q.data.begin(), r.Q.data.end(), Q);
    std::copy(r.m.b.begin(), r.m.b.end(), b);
    if (y) std::copy(r.m.y.begin(), r.m.y.end(), y);
    if (prec) for (size_t i = 0; i < r.m.prec.v.size(); ++i) prec[i] = (unsigned char)r.m.prec.v[i];
    if (firstmatch) std::copy(r.m.firstmatch.begin(), r.m.firstmatch.end(), firstmatch);
}
int cspo_fit_quadratic_native(void* box, int skip, double* c, double* g, double* Q, double* b, double* y, unsigned char* prec,
                              int* firstmatch) {
    CSPO_TRY
    NativeFit r = fit_quadratic_native(kValue, (Box*)box, skip);
    copy_native(r, c, g, Q, b, y, prec, firstmatch);
    CSPO_CATCH
}
int cspo_fit_quadratic(void* box, int skip, double* c, double* gc, double* Q, double* b) {
    CSPO_TRY
    NativeFit r = fit_quadratic_native(kValue, (Box*)box, skip);
    std::vector<double> g = canonicalize(r.m.g, r.Q, r.m.b, r.m.y, r.m.prec);
    *c = r.m.c;
    std::copy(g.begin(), g.end(), gc);
    std::copy(r.Q.data.begin(), r.Q.data.end(), Q);
    std::copy(r.m.b.begin(), r.m.b.end(), b);
    CSPO_CATCH
}
// Q (n*n, SymmetricArray .data layout) is in/out, as in the reference (coefficients_p output goes in).
int cspo_fit_quadratic_gdiag(void* box, int skip, double* Q, double* c, double* g, double* b) {
    CSPO_TRY
    Box* bx = (Box*)box;
    int n = ndims(bx);
    SymArray S(n, 2, 0.0);
    std::copy(Q, Q + (size_t)n * n, S.data.begin());
    std::vector<double> bb = position(bx), gg;
    fit_quadratic_gdiag(kValue, S, bx, bb, skip, *c, gg);
    std::copy(S.data.begin(), S.data.end(), Q);
    std::copy(gg.begin(), gg.end(), g);
    std::copy(bb.begin(), bb.end(), b);
    CSPO_CATCH
}

static Prec mkprec(int n, const unsigned char* prec) {
    Prec P(n);
    for (size_t i = 0; i < (size_t)n * n; ++i) P.v[i] = (char)prec[i];
    return P;
}
static SymArray mksym(int n, const double* Q) {
    SymArray S(n, 2, 0.0);
    std::copy(Q, Q + (size_t)n * n, S.data.begin());
    return S;
}
double cspo_modelvalue_native(int n, double c, const double* g, const double* Q, const double* b, const double* y,
                              const unsigned char* prec, const double* x) {
    return modelvalue_native(c, g, mksym(n, Q), b, y, mkprec(n, prec), x, n);
}
double cspo_modelvalue_chi(int n, double c, const double* g, const double* Q, const double* b, const double* y,
                           const unsigned char* prec, const double* x) {
    return modelvalue_chi(c, g, mksym(n, Q), b, y, mkprec(n, prec), x, n);
}
double cspo_modelvalue_canonical(int n, double c, const double* g, const double* Q, const double* b, const double* x) {
    return modelvalue_canonical(c, g, mksym(n, Q), b, x, n);
}
double cspo_Qcoef_value(int n, int i, int j, const double* x, const double* b, const double* y, const unsigned char* prec) {
    return Qcoef_value(i, j, x, b, y, mkprec(n, prec));
}
double cspo_Qcoef_lineq(int i, int k, double dxi, double xk, double bk, double yk, int precik) {
    return Qcoef_lineq(i, k, dxi, xk, bk, yk, precik != 0);
}
double cspo_chi(int n, int j, int i, const double* b, const double* y, const unsigned char* prec) {
    return chi(j, i, b, y, mkprec(n, prec));
}

// ---- flat export (preorder), computed with the oracle's own walks: used to cross-check the product's flattener
// and to feed oracle-grown trees to the CUDA library in tests.  Arrays as in include/csp_b200.h (csp_tree_desc).
int cspo_tree_counts(void* root, int* n_nodes, int* n_internal, int* n_leaves) {
    CSPO_TRY
    int a = 0, b = 0, c = 0;
    Box* r = (Box*)root;
    for (auto it = box_iterate(r); it.ok; it = box_iterate(r, it.state)) {
        a++;
        if (!isleaf(it.item)) b++;
        else if (!isfake(it.item)) c++;
    }
    *n_nodes = a; *n_internal = b; *n_leaves = c;
    CSPO_CATCH
}
int cspo_export_flat(void* root, int* parent, unsigned char* childindex, int* internal_id, int* leaf_id, double* val,
                     int* inode, int* dims, double* xs, int* child, double* pos) {
    CSPO_TRY
    Box* r = (Box*)root;
    index_tree(r);
    int n = ndims(r), p = r->p, C = 1 << p;
    int I = 0;
    for (auto it = box_iterate(r); it.ok; it = box_iterate(r, it.state)) {
        Box* b = it.item;
        int id = b->preorder_id;
        parent[id] = b->parent->preorder_id;
        childindex[id] = b->childindex;
        leaf_id[id] = b->leaf_rank;
        val[id] = value(b);
        if (isleaf(b)) { internal_id[id] = -1; continue; }
        internal_id[id] = I;
        inode[I] = id;
        for (int k = 0; k < p; ++k) { dims[(size_t)I * p + k] = b->split->dims[k]; xs[(size_t)I * p + k] = b->split->xs[k]; }
        for (int k = 0; k < C; ++k) child[(size_t)I * C + k] = getchild(b->split, k)->preorder_id;
        auto x = position(b);
        std::copy(x.begin(), x.end(), pos + (size_t)I * n);
        I++;
    }
    CSPO_CATCH
}

// ---- batched drivers (OpenMP over queries / leaves): the CPU baseline legs of bench.py and the parity checker.
// leaf_out[i] = rank of the leaf in leaves(root) (index_tree must have been called), status 1 = the reference throws.
int cspo_find_leaf_batch(void* root, const double* X, long m, int n, int* leaf_out, unsigned char* status, int nthreads,
                         double* seconds) {
    CSPO_TRY
    Box* r = (Box*)root;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < m; ++i) {
        try {
            Box* leaf = find_leaf_at(r, X + (size_t)i * n, n);
            leaf_out[i] = leaf->leaf_rank;
            if (status) status[i] = 0;
        } catch (const JlError&) {
            leaf_out[i] = -1;
            if (status) status[i] = 1;
        }
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

// Fit every listed box (pointers from cspo_leaves / cspo_collect).  Outputs: c[nb], g[nb*n] (canonical gc),
// Q[nb*n*n] (.data layout: upper triangle filled, lower NaN), b[nb*n]; optional native outputs gnat[nb*n], y[nb*n],
// prec[nb*n*n].  status: 0 ok, 1 "no chain found" (cs2.jl:167), 2 AssertionError (cs2.jl:186-187), 3 UndefRefError, 9 other.
int cspo_fit_boxes_batch(void** boxes, long nb, int skip, double* c, double* g, double* Q, double* b, double* gnat, double* y,
                         unsigned char* prec, unsigned char* status, int nthreads, double* seconds) {
    CSPO_TRY
    auto t0 = std::chrono::steady_clock::now();
    const double NaN = std::numeric_limits<double>::quiet_NaN();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long l = 0; l < nb; ++l) {
        Box* box = (Box*)boxes[l];
        int n = ndims(box);
        size_t o1 = (size_t)l * n, o2 = (size_t)l * n * n;
        try {
            NativeFit r = fit_quadratic_native(kValue, box, skip);
            std::vector<double> gc = canonicalize(r.m.g, r.Q, r.m.b, r.m.y, r.m.prec);
            c[l] = r.m.c;
            std::copy(gc.begin(), gc.end(), g + o1);
            std::copy(r.Q.data.begin(), r.Q.data.end(), Q + o2);
            std::copy(r.m.b.begin(), r.m.b.end(), b + o1);
            if (gnat) std::copy(r.m.g.begin(), r.m.g.end(), gnat + o1);
            if (y) std::copy(r.m.y.begin(), r.m.y.end(), y + o1);
            if (prec) for (size_t i = 0; i < (size_t)n * n; ++i) prec[o2 + i] = (unsigned char)r.m.prec.v[i];
            status[l] = 0;
        } catch (const JlError& e) {
            c[l] = NaN;
            std::fill(g + o1, g + o1 + n, NaN);
            std::fill(Q + o2, Q + o2 + (size_t)n * n, NaN);
            std::fill(b + o1, b + o1 + n, NaN);
            if (gnat) std::fill(gnat + o1, gnat + o1 + n, NaN);
            if (y) std::fill(y + o1, y + o1 + n, NaN);
            if (prec) std::fill(prec + o2, prec + o2 + (size_t)n * n, (unsigned char)0);
            std::string msg = e.what();
            status[l] = msg.find("no chain found") != std::string::npos ? 1 : e.kind == ERR_ASSERT ? 2 : e.kind == ERR_UNDEFREF ? 3 : 9;
        }
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

int cspo_coefficients_p_batch(void** boxes, long nb, int skip, double* out, long* visited, int nthreads, double* seconds) {
    CSPO_TRY
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long l = 0; l < nb; ++l) {
        Box* box = (Box*)boxes[l];
        CoefStats st;
        SymArray Cp = coefficients_p(kValue, box, skip, &st);
        std::copy(Cp.data.begin(), Cp.data.end(), out + (size_t)l * Cp.data.size());
        if (visited) visited[l] = st.visited;
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

int cspo_coefficients_p_pattern_batch(void** boxes, long nb, int skip, int npat, const int* pat_i, const int* pat_j, double* out,
                                      long* visited, long* used, int nthreads, double* seconds) {
    CSPO_TRY
    std::vector<std::pair<int, int>> pattern;
    for (int k = 0; k < npat; ++k) pattern.emplace_back(pat_i[k], pat_j[k]);
    auto t0 = std::chrono::steady_clock::now();
    std::string err;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long l = 0; l < nb; ++l) {
        try {
            CoefStats st;
            auto v = coefficients_p_pattern(kValue, (Box*)boxes[l], skip, pattern, &st);
            std::copy(v.begin(), v.end(), out + (size_t)l * npat);
            if (visited) visited[l] = st.visited;
            if (used) used[l] = st.used;
        } catch (const std::exception& e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) throw JlError(ERR_ERROR, err);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

// Q = coefficients_p(box; skip); fit_quadratic_gdiag!(Q, box; skip) for a batch of boxes (layouts as in cspo_fit_boxes_batch)
int cspo_fit_gdiag_batch(void** boxes, long nb, int skip, double* c, double* g, double* Q, double* b, int nthreads) {
    CSPO_TRY
    std::string err;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long l = 0; l < nb; ++l) {
        try {
            Box* bx = (Box*)boxes[l];
            int n = ndims(bx);
            SymArray S = coefficients_p(kValue, bx, skip);
            std::vector<double> bb = position(bx), gg;
            double cc;
            fit_quadratic_gdiag(kValue, S, bx, bb, skip, cc, gg);
            c[l] = cc;
            std::copy(gg.begin(), gg.end(), g + (size_t)l * n);
            std::copy(S.data.begin(), S.data.end(), Q + (size_t)l * n * n);
            std::copy(bb.begin(), bb.end(), b + (size_t)l * n);
        } catch (const std::exception& e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) throw JlError(ERR_ERROR, err);
    CSPO_CATCH
}

// possemidef for a batch of matrices in the SymmetricArray .data layout (nb x n*n, column-major each); out: full matrices
int cspo_possemidef_batch(int n, long nb, const double* Q, double* out) {
    CSPO_TRY
    for (long l = 0; l < nb; ++l) {
        SymArray S = mksym(n, Q + (size_t)l * n * n);
        auto r = possemidef(S);
        std::copy(r.begin(), r.end(), out + (size_t)l * n * n);
    }
    CSPO_CATCH
}

// find_leaf_at + canonical modelvalue with per-leaf models indexed by leaf rank (layouts as in cspo_fit_boxes_batch).
int cspo_locate_eval_batch(void* root, const double* X, long m, int n, const double* c, const double* g, const double* Q,
                           const double* b, int* leaf_out, double* val_out, unsigned char* status, int nthreads, double* seconds) {
    CSPO_TRY
    Box* r = (Box*)root;
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < m; ++i) {
        try {
            const double* x = X + (size_t)i * n;
            Box* leaf = find_leaf_at(r, x, n);
            int l = leaf->leaf_rank;
            leaf_out[i] = l;
            val_out[i] = modelvalue_canonical(c[l], g + (size_t)l * n, Q + (size_t)l * n * n, b + (size_t)l * n, x, n);
            if (status) status[i] = 0;
        } catch (const JlError&) {
            leaf_out[i] = -1;
            val_out[i] = NaN;
            if (status) status[i] = 1;
        }
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

// canonical modelvalue for given (leaf rank, x) pairs -- evaluation-only leg (config C4: models supplied per leaf)
int cspo_eval_batch(int n, const double* c, const double* g, const double* Q, const double* b, const double* X, const int* leaf,
                    long m, double* val_out, int nthreads, double* seconds) {
    CSPO_TRY
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < m; ++i) {
        int l = leaf[i];
        val_out[i] = modelvalue_canonical(c[l], g + (size_t)l * n, Q + (size_t)l * n * n, b + (size_t)l * n, X + (size_t)i * n, n);
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CSPO_CATCH
}

}  // extern "C"
