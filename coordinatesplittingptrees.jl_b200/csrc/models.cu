// models.cu -- device-resident per-leaf canonical models (c, g, Q, b): upload of caller-supplied models, download.
// Record layout: see ModelsView in common.cuh.  Q crosses the C ABI in the SymmetricArray .data layout of the reference
// (src/types.jl:432-454): column-major n x n, values at [i,j] with i <= j.
#include "common.cuh"

extern "C" {

int csp_models_free(csp_models* mo) {
    if (!mo) return CSP_OK;
    {
        DeviceGuard g(mo->device);
        cudaDeviceSynchronize();
        csp_dev_free(mo->data);
        csp_dev_free(mo->status);
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
    const int n = mo->n;
    const size_t S = mo->stride;
    const long long slab = std::max<long long>(1, (256LL << 20) / (long long)(S * 8));
    std::vector<double> stage((size_t)std::min<long long>(slab, mo->n_models) * S);
    const double NaN = __builtin_nan("");
    for (long long off = 0; off < mo->n_models; off += slab) {
        const long long cnt = std::min<long long>(slab, mo->n_models - off);
        CSP_CUDA(cudaMemcpy(stage.data(), mo->data + (size_t)off * S, (size_t)cnt * S * 8, cudaMemcpyDeviceToHost));
        for (long long l = 0; l < cnt; ++l) {
            const double* r = stage.data() + (size_t)l * S;
            const long long L = off + l;
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
    if (status) CSP_CUDA(cudaMemcpy(status, mo->status, (size_t)mo->n_models, cudaMemcpyDeviceToHost));
    return CSP_OK;
}

}  // extern "C"
