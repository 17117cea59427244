/* csp_b200.h -- C ABI of libcsp_b200.so: the B200 (sm_100a) accelerator for the data-parallel hot path of
 * timholy/CoordinateSplittingPTrees.jl (batched find_leaf_at, per-box quadratic model fit, batched model evaluation).
 *
 * The reference is pure Julia and has no FFI of its own; the boundary below is what a Julia `ccall` binding for this
 * path would bind (see INTEGRATION.md for the Julia stub).  Each entry point cites the reference function it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns CSP_OK (0) or a negative CSP_ERR_* code; csp_last_error() returns the message of the last
 *     failure on the calling thread.  No exceptions cross this boundary; the host wrapper re-raises the reference's
 *     exception types from the status codes documented below.
 *   - the caller owns every host buffer; the library owns device memory behind the opaque handles.
 *   - threads: point location and evaluation only read a csp_tree / csp_models and may run concurrently on one handle (from
 *     different host threads and streams); the fit entry points share a per-tree scratch buffer and are NOT reentrant on one
 *     handle; handles are created / freed / re-fitted by one thread at a time; the multi-GPU layer and the timing switch are
 *     process-wide and driven by one thread.
 *   - a query matrix X holds m queries of n doubles each, each query contiguous (Julia: an n x m column-major Matrix).
 *   - `*_dev` variants take DEVICE pointers on the handle's device plus a cudaStream_t (as void*; NULL = default stream),
 *     are asynchronous, and are what a caller with resident data uses.  The plain variants take HOST buffers, move them
 *     with chunked, double-buffered copies and return when the results are in the caller's buffers.
 *   - dimensions are 1-based in every array that holds them (as in the reference); node / leaf ids are 0-based.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with CSP_ERR_NO_DEVICE.
 */
#ifndef CSP_B200_H
#define CSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSP_OK 0
#define CSP_ERR_INVALID (-1)     /* bad argument; DimensionMismatch / ArgumentError on the host side */
#define CSP_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define CSP_ERR_NOMEM (-3)
#define CSP_ERR_UNSUPPORTED (-4) /* degree / dimension outside what the kernels cover (fits need p == 2, models p <= 2) */
#define CSP_ERR_NO_DEVICE (-5)
#define CSP_ERR_NCCL (-6)        /* NCCL could not be loaded, or a collective failed */

/* per-query status written by csp_find_leaf* / csp_locate_eval* */
#define CSP_Q_OK 0
#define CSP_Q_OUTSIDE 1 /* src/tree.jl:357-360 throws ErrorException("... not within ... along dimension ...") */

/* per-box status written by csp_fit_boxes / csp_models_fit */
#define CSP_FIT_OK 0
#define CSP_FIT_NO_CHAIN 1     /* src/cs2.jl:166-167  error("no chain found at ...") */
#define CSP_FIT_CHAIN_ASSERT 2 /* src/cs2.jl:186-187  AssertionError (chain repeats a dimension) */
#define CSP_FIT_CHAIN_UNDEF 3  /* src/cs2.jl:184      UndefRefError (chain runs into a leaf before covering n dims) */

typedef struct csp_tree csp_tree;     /* device-resident flattened CSp-tree */
typedef struct csp_models csp_models; /* device-resident per-leaf models (c, g, Q, b) */

/* Flattened tree, as produced by walking `collect(root)` (preorder, src/tree.jl:393-415).
 * Node ids are preorder ranks; internal ids are preorder ranks among non-leaf boxes; leaf ids are ranks in
 * `leaves(root)` (src/tree.jl:423, fake boxes excluded).  Replaces the Box/Split/Children/World object graph of
 * src/types.jl:7-23,155-192.  Everything derivable (split midpoints, subtree sizes, child values ...) is computed by
 * the library at upload. */
typedef struct csp_tree_desc {
    int32_t n;          /* ndims(root) */
    int32_t p;          /* degree(root): 1..8 (src/types.jl:194-201); fits and models need p <= 2 */
    int32_t n_nodes;    /* length(root) */
    int32_t n_internal; /* number of non-leaf boxes */
    int32_t n_leaves;   /* length(leaves(root)) */
    const double* lower;        /* [n]  world.lower */
    const double* upper;        /* [n]  world.upper */
    /* per node, preorder */
    const int32_t* parent;      /* [n_nodes] node id of box.parent (root: itself) */
    const uint8_t* childindex;  /* [n_nodes] box.childindex (0 = self) */
    const int32_t* internal_id; /* [n_nodes] internal id, or -1 for leaves */
    const int32_t* leaf_id;     /* [n_nodes] leaf id, or -1 for non-leaves and fake leaves (src/tree.jl:430-442) */
    const double* val;          /* [n_nodes] value(box) (src/CoordinateSplittingPTrees.jl:61, src/tree.jl:316-323) */
    /* per internal node, preorder */
    const int32_t* inode;       /* [n_internal] node id */
    const int32_t* dims;        /* [n_internal*p] split.dims, 1-based, > n = fictive */
    const double* xs;           /* [n_internal*p] split.xs */
    const int32_t* child;       /* [n_internal*2^p] node ids of getchild(split, 0..2^p-1) (src/tree.jl:325) */
    const double* pos;          /* [n_internal*n] position(box) (src/tree.jl:71,103-129) */
} csp_tree_desc;

int csp_version(void);
const char* csp_last_error(void);
int csp_device_count(int* count);
int csp_set_device(int device); /* device used by subsequent uploads on this thread (default 0) */

/* ---- tree residency ------------------------------------------------------------------------------------------- */
int csp_tree_upload(const csp_tree_desc* desc, csp_tree** out);
/* The same tree as one relocatable byte blob (the wire / on-disk format: header + the desc arrays, little endian),
 * so that it can be broadcast across GPUs or saved.  `blob` for csp_tree_upload_blob may be a host or a device pointer. */
int csp_tree_blob_size(const csp_tree_desc* desc, size_t* bytes);
int csp_tree_pack(const csp_tree_desc* desc, void* blob, size_t bytes);
int csp_tree_upload_blob(const void* blob, size_t bytes, csp_tree** out);
/* ---- incremental residency: a tree that grows on the device ------------------------------------------------
 * addpoint! (src/CoordinateSplittingPTrees.jl:94-142) appends ceil(n/p) splits per call and shifts the preorder ids of everything
 * behind them; re-flattening and re-uploading the whole tree per call is O(tree).  A LIVE tree is fed the SPLIT LOG instead -- the
 * splits in creation order, where ids never change: box 0 is the root and the s-th split of the tree (0-based) creates the boxes
 * 1 + s*2^p + (0 .. 2^p-1), child index = offset.  An append sends only the new splits; the library re-derives the preorder
 * flattening and everything else on the device.  All other entry points work on a live tree as on an uploaded one (node and
 * leaf ids are the preorder ids of the CURRENT version; csp_tree_live_maps translates box ids). */
int csp_tree_create_live(int32_t n, int32_t p, const double* lower, const double* upper, const double* root_position,
                         double root_value, csp_tree** out);
/* The splits made since the last call, in creation order: entry k divides box split_box[k] along dims[k*p ..] (1-based, > n
 * fictive) at xs[k*p ..]; child_val[k*2^p + c] = value(child c) (child 0 = split.self repeats the divided box's value). */
int csp_tree_append_splits(csp_tree* tree, int64_t nsplits, const int32_t* split_box, const int32_t* dims, const double* xs,
                           const double* child_val);
/* node_of_box[b] = current preorder node id of box b, box_of_node[v] = box id of node v (n_nodes entries each; either may be NULL). */
int csp_tree_live_maps(const csp_tree* tree, int32_t* node_of_box, int32_t* box_of_node);

/* Diagnostic: one of the arrays the library derives on the device at upload, copied back to the host (tests compare the device
 * builder with the host-side construction, CSP_UPLOAD_HOST=1).  which: 0 node records, 1 descent records, 2 descent leaf table,
 * 3 split-box coordinates, 4 child values, 5 leaf nodes, 6 path offsets, 7 ancestor paths, 8 fit units.  *bytes = its size; out may
 * be NULL to query it. */
int csp_tree_export_derived(const csp_tree* tree, int which, void* out, size_t cap, size_t* bytes);
/* info[0..7] = n, p, n_nodes, n_internal, n_leaves, device, max leaf depth, bytes resident (KiB) */
int csp_tree_info(const csp_tree* tree, int64_t info[8]);
int csp_tree_free(csp_tree* tree);

/* ---- K1: batched point location; replaces find_leaf_at (src/tree.jl:351-388) ------------------------------------ */
/* leaf[i] = leaf id of the leaf containing X[:,i], or -1 with status[i] = CSP_Q_OUTSIDE.  status may be NULL. */
int csp_find_leaf(const csp_tree* tree, const double* X, int64_t m, int32_t* leaf, uint8_t* status);
int csp_find_leaf_dev(const csp_tree* tree, const double* dX, int64_t m, int32_t* dleaf, uint8_t* dstatus, void* stream);
/* find_leaf_at(box, x) with ANY box of the tree as the starting point (the reference accepts one, and addpoint! calls it that way:
 * src/CoordinateSplittingPTrees.jl:106-108): the general bounds check of src/tree.jl:357-360 against boxbounds(box, i)
 * (src/tree.jl:146-171) -- lower <= x < upper, the upper edge allowed only where it is the world's -- then the descent from
 * that box.  node_id = 0 is the root. */
int csp_find_leaf_from(const csp_tree* tree, int32_t node_id, const double* X, int64_t m, int32_t* leaf, uint8_t* status);
int csp_find_leaf_from_dev(const csp_tree* tree, int32_t node_id, const double* dX, int64_t m, int32_t* dleaf, uint8_t* dstatus,
                           void* stream);
/* boxbounds(box) (src/tree.jl:136-171) of any box, computed on the device: lower[n], upper[n] (host arrays). */
int csp_boxbounds(const csp_tree* tree, int32_t node_id, double* lower, double* upper);
/* minimum(box) / maximum(box) / extrema(box) (src/CoordinateSplittingPTrees.jl:64-68: isless on value over leaves(box)): ids
 * and values of the first minimal and the first maximal real leaf under node_id (NaN sorts last, -0.0 before 0.0, as isless does).
 * A device reduction over the resident values: it can follow an upload or a fit without the values visiting the host.
 * Any output may be NULL. */
int csp_tree_extrema(const csp_tree* tree, int32_t node_id, int32_t* leaf_min, double* val_min, int32_t* leaf_max, double* val_max);

/* ---- K2: batched model fits ------------------------------------------------------------------------------------- */
/* fit_quadratic(value, box; skip) for each listed box (src/cs2.jl:89-93: fit_poly1 :162-205, coefficients_p
 * src/polynomials.jl:182-230, fit_poly2_diag! src/cs2.jl:208-274, canonicalize :132-140).  node_ids == NULL means
 * "every leaf, in leaves(root) order" (then nb must equal n_leaves).  Outputs (any may be NULL), box-major:
 *   c[nb], g[nb*n] canonical gradient, Q[nb*n*n] in the layout of SymmetricArray .data (column-major, values at
 *   [i,j] with i<=j, the strictly lower triangle keeps the NaN fill; src/types.jl:432-454, src/polynomials.jl:232-233),
 *   b[nb*n]; native-form extras of fit_quadratic_native (src/cs2.jl:115-121): g_native[nb*n], y[nb*n],
 *   prec[nb*n*n] (column-major Bool matrix, prec[i,j] <=> i was split before-or-with j), firstmatch[nb*n] (src/cs2.jl:173,
 *   191-193: the 1-based dimension split together with each dimension in the chain, n+1 for a fictive partner); status[nb] CSP_FIT_*.
 * Boxes whose status is not CSP_FIT_OK get NaN outputs.  Requires p == 2. */
int csp_fit_boxes(const csp_tree* tree, const int32_t* node_ids, int64_t nb, int32_t skip, double* c, double* g, double* Q,
                  double* b, double* g_native, double* y, uint8_t* prec, int32_t* firstmatch, uint8_t* status);
/* coefficients_p(value, box; skip) (src/polynomials.jl:127-230), dense output Cp[nb*n^p] in the column-major .data layout of the
 * p-dimensional SymmetricArray (values at sorted index tuples, NaN elsewhere): p == 2 gives the off-diagonal Hessian entries
 * (diagonal NaN), p == 1 the gradient vector, 3 <= p <= 8 the p-th order mixed coefficients (test/runtests.jl:593-648).
 * visited[nb] (may be NULL) = number of splits the traversal examined; used[nb] (may be NULL) = number of splits that supplied a
 * coefficient -- the reference invokes its callback 2^p times for each of them (src/polynomials.jl:222, counts at
 * test/runtests.jl:652-700). */
int csp_coefficients_p(const csp_tree* tree, const int32_t* node_ids, int64_t nb, int32_t skip, double* Cp, int64_t* visited,
                       int64_t* used);

/* Q = coefficients_p(value, box; skip) followed by fit_quadratic_gdiag!(value, Q, box; skip) (src/cs2.jl:332-384): the chain-free
 * fit -- gradient and diagonal from two 1-d difference equations per dimension, b = position(box), c = value(box).  It never
 * fails with "no chain found"; undetermined entries are NaN.  Outputs as in csp_fit_boxes (g is already canonical).  Requires p == 2. */
int csp_fit_gdiag(const csp_tree* tree, const int32_t* node_ids, int64_t nb, int32_t skip, double* c, double* g, double* Q, double* b);
/* coefficients_p!(value, Cp, box; skip) into a SPARSE symmetric output (src/polynomials.jl:161-230 with nnz_offdiag_sym :244-259,
 * fillnz! :261-271): only the listed above-diagonal positions (pat_i[k] < pat_j[k], 1-based) are filled and the traversal stops as soon
 * as all of them are; a SymTridiagonal output is the pattern (i, i+1).  vals[nb*npat] in pattern order (NaN where no split supplies
 * the pair).  Requires p == 2. */
int csp_coefficients_p_sparse(const csp_tree* tree, const int32_t* node_ids, int64_t nb, int32_t skip, int32_t npat,
                              const int32_t* pat_i, const int32_t* pat_j, double* vals, int64_t* visited, int64_t* used);

/* ---- per-leaf models resident on the device ---------------------------------------------------------------------- */
/* Fit every leaf and keep the canonical models (c, g, Q, b) on the device for evaluation.  p == 2: fit_quadratic;
 * p == 1: the degree-1 model c = value(leaf), g = coefficients_p(leaf), b = position(leaf). */
int csp_models_fit(const csp_tree* tree, int32_t skip, csp_models** out);
/* Re-fit into an existing handle obtained from csp_models_fit on the same tree (asynchronous on `stream`; no allocation):
 * what a caller that re-fits every iteration uses. */
int csp_models_refit(const csp_tree* tree, int32_t skip, csp_models* models, void* stream);
/* An unfitted handle for this tree (records zeroed): the target of csp_models_refit / csp_models_refit_range. */
int csp_models_alloc(const csp_tree* tree, csp_models** out);
/* Re-fit only the leaves [leaf_lo, leaf_hi) (ids in leaves(root) order) into an existing handle; the other records are left
 * untouched.  This is the per-GPU piece of a leaf-sharded fit (csp_shard_fit).  Asynchronous on `stream`. */
int csp_models_refit_range(const csp_tree* tree, int32_t skip, csp_models* models, int64_t leaf_lo, int64_t leaf_hi, void* stream);
/* The records of the listed leaves only (nb ids), outputs as in csp_models_download (any may be NULL). */
int csp_models_download_rows(const csp_models* models, const int32_t* leaf_ids, int64_t nb, double* c, double* g, double* Q,
                             double* b, uint8_t* status);
/* Models supplied by the caller (host arrays, layouts as in csp_fit_boxes; Q may be NULL for degree-1 models). */
int csp_models_upload(int32_t n, int64_t n_models, const double* c, const double* g, const double* Q, const double* b,
                      csp_models** out);
/* possemidef(Q) (src/posdef.jl:41-68) applied in place to the Q of every resident model: off-diagonals kept, diagonals raised as
 * needed to make Q positive semi-definite (NaN off-diagonals count as 0).  Asynchronous on `stream`. */
int csp_models_possemidef(csp_models* models, void* stream);
int csp_models_download(const csp_models* models, double* c, double* g, double* Q, double* b, uint8_t* status);
/* info[0..3] = n, n_models, degree (1|2), device */
int csp_models_info(const csp_models* models, int64_t info[4]);
int csp_models_free(csp_models* models);

/* ---- K3: batched evaluation; replaces modelvalue(c, g, Q, b, x) (src/cs2.jl:155-159) ----------------------------- */
/* val[i] = c_l + g_l'(x_i - b_l) + (x_i - b_l)' Q_l (x_i - b_l) / 2 with l = leaf[i]; a leaf id outside
 * [0, n_models) -- negative (out-of-world marker) or too large -- gives NaN on every evaluation path. */
int csp_eval(const csp_models* models, const double* X, const int32_t* leaf, int64_t m, double* val);
int csp_eval_dev(const csp_models* models, const double* dX, const int32_t* dleaf, int64_t m, double* dval, void* stream);
/* find_leaf_at followed by modelvalue of that leaf's model, fused.  leaf / status may be NULL. */
int csp_locate_eval(const csp_tree* tree, const csp_models* models, const double* X, int64_t m, int32_t* leaf, double* val,
                    uint8_t* status);
int csp_locate_eval_dev(const csp_tree* tree, const csp_models* models, const double* dX, int64_t m, int32_t* dleaf,
                        double* dval, uint8_t* dstatus, void* stream);

/* Native-form evaluation of ONE CS2-native model (c, g, Q, b, y, prec) as returned by fit_quadratic_native, at m points:
 * modelvalue(c, g, Q, b, y, prec, x) (src/cs2.jl:142-153 with Qcoef_value :302-312), in the reference's loop order.
 * Q: n*n in the .data layout, prec: n*n column-major bytes.  Host buffers. */
int csp_modelvalue_native(int32_t n, double c, const double* g, const double* Q, const double* b, const double* y,
                          const uint8_t* prec, const double* X, int64_t m, double* val);

/* ---- multi-GPU: the tree replicated on every GPU, queries and leaf fits sharded, results gathered ------
 * The reference is single-threaded and has no counterpart; the partition is BASELINE.json's (configs[2], configs[4]): the
 * flattened tree is NCCL-broadcast once, query batches shard by contiguous ranges and leaf fits by contiguous leaf-id ranges with
 * no exchange between ranks except the all-gather of the fitted records, and (leaf id, value) results are gathered to one rank.
 * NCCL (libnccl.so.2) is loaded when the first of these functions is called.
 *
 * Two ways to form the clique.  Single process: csp_init makes the listed GPUs ranks 0..ndev-1 (ncclCommInitAll), all driven by
 * the calling thread -- what a Julia host uses.  One process per GPU (torchrun): rank 0 obtains 128 bytes from
 * csp_comm_unique_id, hands them to the other ranks by any side channel, and every rank calls csp_init_rank (ncclCommInitRank).
 * Every csp_shard_* call is collective: each process calls it with the same arguments; per-rank arrays (`dX`, `dleaf`, `streams`
 * ...) have one entry per LOCAL rank (ndev in the first mode, 1 in the second).  `streams` may be NULL (the library's own
 * non-blocking stream per rank) or hold one cudaStream_t per local rank (a NULL entry is that device's default stream, as in the
 * *_dev calls); the call's work, communication included, is ordered on those streams and is asynchronous. */
int csp_init(int ndev, const int* devs); /* devs == NULL: devices 0..ndev-1 */
int csp_comm_unique_id(void* id128);
int csp_init_rank(int world, int rank, int device, const void* id128);
int csp_shutdown(void);
/* info[0..3] = world size, local ranks, first local rank, NCCL version code (0 before initialisation) */
int csp_comm_info(int64_t info[4]);

typedef struct csp_shard_tree csp_shard_tree;     /* one csp_tree replica per local rank */
typedef struct csp_shard_models csp_shard_models; /* one csp_models replica per local rank */

/* C1: broadcast the packed tree (csp_tree_pack) from rank `root` and upload it on every GPU straight from the device buffer the
 * broadcast landed in.  `blob`/`bytes` are read only by the process that owns rank `root`. */
int csp_shard_tree_broadcast(const void* blob, size_t bytes, int root, csp_shard_tree** out);
int csp_shard_tree_get(const csp_shard_tree* st, int local_index, csp_tree** tree); /* borrowed: do not free */
int csp_shard_tree_free(csp_shard_tree* st);
/* Leaf fits sharded by contiguous leaf-id range over the ranks (csp_models_refit_range), then the fitted records are all-gathered
 * so that every GPU holds every model.  *models == NULL allocates the replicas, otherwise they are re-fitted in place.
 * The all-gather runs on the COPY ENGINES when the GPUs can map each other's memory (peer access / CUDA IPC; the replicas are
 * plain cudaMalloc memory then): behind its own fit, every rank pushes its record range into every other replica by
 * device-to-device DMA over NVLink on a communication stream, between two one-word barriers.  The call does NOT make the
 * callers' streams wait for it: each csp_models carries a `ready` event that every consumer of the records (evaluation,
 * download, possemidef) waits for, so the descent of the next batch overlaps the transfer.  Otherwise (CSP_ALLGATHER=nccl, or
 * no peer mapping): one in-place ncclBroadcast per owner range, grouped, on the callers' streams. */
int csp_shard_fit(const csp_shard_tree* st, int32_t skip, csp_shard_models** models, void* const* streams);
int csp_shard_models_get(const csp_shard_models* sm, int local_index, csp_models** models); /* borrowed */
int csp_shard_models_free(csp_shard_models* sm);
/* Resident shards: rank r holds counts[r] queries (counts has WORLD entries, identical on every process); local rank i reads
 * dX[i] and writes dleaf[i] / dval[i] / dstatus[i] (dstatus may be NULL or hold NULLs).  With gather_root >= 0 the (leaf, value)
 * pairs of all ranks are also gathered, in rank order, into dleaf_all / dval_all on rank gather_root's GPU (C2: grouped
 * ncclSend/ncclRecv, pipelined chunk by chunk behind the kernels of the next chunk; chunk = queries per pipeline step, 0 = default).
 * dleaf_all / dval_all are read only by the process that owns gather_root. */
int csp_shard_locate_eval_dev(const csp_shard_tree* st, const csp_shard_models* sm, const double* const* dX, const int64_t* counts,
                              int32_t* const* dleaf, double* const* dval, uint8_t* const* dstatus, int gather_root,
                              int32_t* dleaf_all, double* dval_all, int64_t gather_base, int64_t chunk, void* const* streams);
/* The gather alone (results already computed): what csp_shard_locate_eval_dev does after its kernels. */
int csp_shard_gather_dev(const int32_t* const* dleaf, const double* const* dval, const int64_t* counts, int gather_root,
                         int32_t* dleaf_all, double* dval_all, int64_t gather_base, void* const* streams);
/* By default every csp_shard_locate_eval_dev call ends with the callers' streams waiting for its last gather step.  With
 * csp_shard_defer_join(1) they do not: consecutive calls then overlap the previous call's last transfer with their own kernels
 * (the calls must write disjoint result ranges), and csp_shard_join makes the streams wait for every gather issued so far. */
int csp_shard_defer_join(int on);
int csp_shard_join(void* const* streams);
/* Gather buffers owned by the library, `capacity` results on rank `root`'s GPU (collective; *dleaf_all / *dval_all are set on the
 * process that owns `root`, NULL elsewhere).  While they exist, a gather to that root ignores the dleaf_all / dval_all arguments
 * and writes the call's block at element offset `gather_base` of these buffers (every process passes the same gather_base),
 * moving it with the GPUs' COPY ENGINES -- device-to-device DMA over NVLink straight into the root's memory (peer access in
 * single-process mode, a CUDA IPC mapping with one process per GPU) -- so that no SM is taken from the kernels the transfer
 * overlaps; NCCL only carries one completion word per call.  Without them the gather runs as NCCL send / receive kernels.
 * A rank may compute straight into its own block of the buffers (then nothing is copied for it).  CSP_ERR_UNSUPPORTED when the
 * GPUs cannot map each other's memory. */
int csp_shard_gather_alloc(int64_t capacity, int root, int32_t** dleaf_all, double** dval_all);
int csp_shard_gather_free(void);
/* Host buffers: the calling process's m queries are split over its local GPUs (contiguous ranges, one host thread and one copy /
 * compute pipeline per GPU) and the results land in the caller's arrays -- in single-process mode that is the whole batch
 * gathered on the host.  Returns when the results are complete. */
int csp_shard_locate_eval(const csp_shard_tree* st, const csp_shard_models* sm, const double* X, int64_t m, int32_t* leaf,
                          double* val, uint8_t* status);
/* Wait for the library streams of every local rank. */
int csp_shard_synchronize(void);
/* Device milliseconds of the last collective of each kind, measured with CUDA events on local rank 0's stream (synchronizes):
 * ms[0] tree blob host->device on the root, ms[1] tree ncclBroadcast, ms[2] per-GPU upload + derived arrays (host wall clock),
 * ms[3] sharded fit kernels, ms[4] all-gather of the records, ms[5] gather of the results (csp_shard_gather_dev). */
int csp_shard_timing(double ms[6]);

/* ---- instrumentation -------------------------------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process since load (bench.py reports it as gpu_launches). */
int64_t csp_kernel_launches(void);
/* Per-kernel device times, measured with CUDA events recorded on the launching stream around every kernel launch while
 * enabled.  csp_timing_read synchronizes the device, returns the accumulated milliseconds and launch counts per kernel
 * class and clears them.  Classes: 0 locate (K1, +histogram), 1 fused locate+evaluate, 2 scan, 3 scatter,
 * 4 binned evaluation, 5 evaluate-only, 6 fit (K2); ms and launches have 7 entries. */
int csp_timing_enable(int on);
int csp_timing_read(double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* CSP_B200_H */
