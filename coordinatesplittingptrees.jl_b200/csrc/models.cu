// models.cu -- device-resident per-leaf canonical models (c, g, Q, b): upload of caller-supplied models, download.
// Record layout: see ModelsView in common.cuh.  Q crosses the C ABI in the SymmetricArray .data layout of the reference
// (src/types.jl:432-454): column-major n x n, values at [i,j] with i <= j.
#include "common.cuh"

// record -> (c, g, Q in the .data layout, b) for cnt records of a staged slab; output row of record l is L0 + l
static void unpack_records(const csp_models* mo, const double* stage, long long cnt, long long L0, double* c, double* g, double* Q, double* b) {
    const int n = mo->n;
    const size_t S = mo->stride;
    const double NaN = __builtin_nan("");
    for (long long l = 0; l < cnt; ++l) {
        const double* r = stage + (size_t)l * S;
        const long long L = L0 + l;
        for (int i = 0; i < n; ++i)
            if (b) b[L * n + i] = r[i];
        if (mo->degree == 2) {
            const double* A = r + n;
            if (Q) {
                double* q = Q + (size_t)L * n * n;
                for (size_t e = 0; e < (size_t)n * n; ++e) q[e] = NaN;  // the SymmetricArray fill value below the diagonal
                for (int i = 0; i < n; ++i)
                    for (int j = i; j < n; ++j) q[(size_t)j * n + i] = (i == j ? 2.0 : 1.0) * A[csp_a_index(n, i, j)];
            }
            for (int i = 0; i < n; ++i)
                if (g) g[L * n + i] = A[csp_a_index(n, i, n)];
            if (c) c[L] = A[csp_a_index(n, n, n)];
        } else {
            for (int i = 0; i < n; ++i)
                if (g) g[L * n + i] = r[n + i];
            if (c) c[L] = r[2 * n];
        }
    }
}

__global__ void __launch_bounds__(128) k_gather_rows(const double* __restrict__ data, const uint8_t* __restrict__ status, int stride,
                                                     const int* __restrict__ ids, long long nb, double* __restrict__ out,
                                                     uint8_t* __restrict__ st_out) {
    for (long long r = blockIdx.x; r < nb; r += gridDim.x) {
        const double* src = data + (size_t)ids[r] * stride;
        double* dst = out + (size_t)r * stride;
        for (int e = threadIdx.x; e < stride; e += blockDim.x) dst[e] = src[e];
        if (threadIdx.x == 0) st_out[r] = status[ids[r]];
    }
}

extern "C" {

int csp_models_free(csp_models* mo) {
    if (!mo) return CSP_OK;
    {
        DeviceGuard g(mo->device);
        cudaDeviceSynchronize();
        if (mo->plain) { cudaFree(mo->data); cudaFree(mo->status); }
        else { csp_dev_free(mo->data); csp_dev_free(mo->status); }
        if (mo->ready) cudaEventDestroy(mo->ready);
    }
    delete mo;
    return CSP_OK;
}

int csp_models_info(const csp_models* mo, int64_t info[4]) {
    if (!mo || !info) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    info[0] = mo->n; info[1] = mo->n_models; info[2] = mo->degree; info[3] = mo->device;
    return CSP_OK;
}

int csp_models_upload(int32_t n, int64_t n_models, const double* c, const double* g, const double* Q, const double* b,
                      csp_models** out) {
    if (!out) return csp_set_error(CSP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n < 1 || n_models < 1 || !c || !g || !b) return csp_set_error(CSP_ERR_INVALID, "bad model arrays");
    CSP_CHECK(csp_require_device());
    csp_models* mo = new csp_models();
    mo->device = csp_current_device();
    mo->n = n; mo->degree = Q ? 2 : 1; mo->n_models = n_models;
    mo->stride = csp_model_stride(n, mo->degree);
    DeviceGuard dg(mo->device);
    const size_t S = mo->stride;
    // pack on the host in slabs (bounded staging memory), then copy
    const long long slab = std::max<long long>(1, (256LL << 20) / (long long)(S * 8));
    std::vector<double> stage((size_t)std::min<long long>(slab, n_models) * S);
    auto fail = [&](int rc) { csp_models_free(mo); return rc; };
    if (csp_dev_alloc((void**)&mo->data, (size_t)n_models * S * 8) != CSP_OK || csp_dev_alloc((void**)&mo->status, (size_t)n_models) != CSP_OK)
        return fail(csp_set_error(CSP_ERR_NOMEM, "cannot allocate %lld model records", (long long)n_models));
    if (cudaMemset(mo->status, 0, (size_t)n_models) != cudaSuccess) return fail(csp_set_error(CSP_ERR_CUDA, "memset failed"));
    for (long long off = 0; off < n_models; off += slab) {
        const long long cnt = std::min<long long>(slab, n_models - off);
        std::fill(stage.begin(), stage.begin() + (size_t)cnt * S, 0.0);
        for (long long l = 0; l < cnt; ++l) {
            double* r = stage.data() + (size_t)l * S;
            const long long L = off + l;
            for (int i = 0; i < n; ++i) r[i] = b[L * n + i];
            if (Q) {
                double* A = r + n;
                const double* q = Q + (size_t)L * n * n;
                for (int i = 0; i < n; ++i) {
                    for (int j = i; j < n; ++j) A[csp_a_index(n, i, j)] = (i == j ? 0.5 : 1.0) * q[(size_t)j * n + i];  // .data[i,j], i <= j
                    A[csp_a_index(n, i, n)] = g[L * n + i];
                }
                A[csp_a_index(n, n, n)] = c[L];
            } else {
                for (int i = 0; i < n; ++i) r[n + i] = g[L * n + i];
                r[2 * n] = c[L];
            }
        }
        if (cudaMemcpy(mo->data + (size_t)off * S, stage.data(), (size_t)cnt * S * 8, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(csp_set_error(CSP_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    *out = mo;
    return CSP_OK;
}

int csp_models_download(const csp_models* mo, double* c, double* g, double* Q, double* b, uint8_t* status) {
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    DeviceGuard dg(mo->device);
    if (mo->ready) CSP_CUDA(cudaEventSynchronize(mo->ready));
    const size_t S = mo->stride;
    const long long slab = std::max<long long>(1, (256LL << 20) / (long long)(S * 8));
    std::vector<double> stage((size_t)std::min<long long>(slab, mo->n_models) * S);
    for (long long off = 0; off < mo->n_models; off += slab) {
        const long long cnt = std::min<long long>(slab, mo->n_models - off);
        CSP_CUDA(cudaMemcpy(stage.data(), mo->data + (size_t)off * S, (size_t)cnt * S * 8, cudaMemcpyDeviceToHost));
        unpack_records(mo, stage.data(), cnt, off, c, g, Q, b);
    }
    if (status) CSP_CUDA(cudaMemcpy(status, mo->status, (size_t)mo->n_models, cudaMemcpyDeviceToHost));
    return CSP_OK;
}

int csp_models_download_rows(const csp_models* mo, const int32_t* leaf_ids, int64_t nb, double* c, double* g, double* Q, double* b,
                             uint8_t* status) {
    if (!mo || (nb > 0 && !leaf_ids) || nb < 0) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    for (int64_t i = 0; i < nb; ++i)
        if (leaf_ids[i] < 0 || leaf_ids[i] >= mo->n_models) return csp_set_error(CSP_ERR_INVALID, "leaf id %d out of range at %lld", leaf_ids[i], (long long)i);
    if (nb == 0) return CSP_OK;
    DeviceGuard dg(mo->device);
    if (mo->ready) CSP_CUDA(cudaEventSynchronize(mo->ready));
    const size_t S = mo->stride;
    const long long slab = std::max<long long>(1, (128LL << 20) / (long long)(S * 8));
    const long long cap = std::min<long long>(slab, nb);
    std::vector<double> stage((size_t)cap * S);
    std::vector<uint8_t> hst((size_t)cap);
    void *dids = nullptr, *drows = nullptr, *dst = nullptr;
    CSP_CHECK(csp_dev_alloc(&dids, (size_t)cap * 4));
    int rc = csp_dev_alloc(&drows, (size_t)cap * S * 8);
    if (rc == CSP_OK) rc = csp_dev_alloc(&dst, (size_t)cap);
    for (long long off = 0; off < nb && rc == CSP_OK; off += cap) {
        const long long cnt = std::min<long long>(cap, nb - off);
        cudaError_t e = cudaMemcpy(dids, leaf_ids + off, (size_t)cnt * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            k_gather_rows<<<(unsigned)std::min<long long>(cnt, 148 * 16), 128>>>(mo->data, mo->status, (int)S, (const int*)dids, cnt, (double*)drows, (uint8_t*)dst);
            CSP_LAUNCHED();
            e = cudaMemcpy(stage.data(), drows, (size_t)cnt * S * 8, cudaMemcpyDeviceToHost);
        }
        if (e == cudaSuccess) e = cudaMemcpy(hst.data(), dst, (size_t)cnt, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = csp_set_error(CSP_ERR_CUDA, "row download failed: %s", cudaGetErrorString(e)); break; }
        unpack_records(mo, stage.data(), cnt, 0, c ? c + off : nullptr, g ? g + off * mo->n : nullptr,
                       Q ? Q + (size_t)off * mo->n * mo->n : nullptr, b ? b + off * mo->n : nullptr);
        if (status) std::copy(hst.begin(), hst.begin() + cnt, status + off);
    }
    csp_dev_free(dids); csp_dev_free(drows); csp_dev_free(dst);
    return rc;
}

}  // extern "C"
