#include "rollout_wt.cuh"
using namespace pime;
extern "C" int pime_wt_rollout_f32(const pime_wt_config *cfg, int64_t n, const pime_wt_state *st, const pime_rollout_args *args,
                                   void *stream) {
    return wt_rollout_impl<float>(cfg, n, st, args, stream);
}

// Host-buffer entry: H2D of the per-env state, fused rollout, D2H of ep_return (+ final state).
extern "C" int pime_wt_rollout_host_f32(const pime_wt_config *cfg, int64_t n, const pime_wt_state *h, const pime_wt_state *d,
                                        const pime_rollout_args *args, float *ep_return_host, void *stream) {
    PIME_REQUIRE(cfg && h && d && args, "null pointer");
    PIME_REQUIRE(d->ep_return, "device ep_return scratch is required");
    if (int rc = require_device()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t b = (size_t)n * sizeof(float);
    void *const hs[7] = {h->h1, h->h2, h->r, h->I, h->a1, h->a2, h->Kp};
    void *const ds[7] = {d->h1, d->h2, d->r, d->I, d->a1, d->a2, d->Kp};
    for (int j = 0; j < 7; ++j)
        if (hs[j] && ds[j]) PIME_CUDA(cudaMemcpyAsync(ds[j], hs[j], b, cudaMemcpyHostToDevice, s));
    if (h->t) PIME_CUDA(cudaMemcpyAsync(d->t, h->t, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (h->episode) PIME_CUDA(cudaMemcpyAsync(d->episode, h->episode, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    PIME_CUDA(cudaMemsetAsync(d->ep_return, 0, b, s));
    if (int rc = wt_rollout_impl<float>(cfg, n, d, args, stream)) return rc;
    if (ep_return_host) PIME_CUDA(cudaMemcpyAsync(ep_return_host, d->ep_return, b, cudaMemcpyDeviceToHost, s));
    for (int j = 0; j < 4; ++j)
        if (hs[j] && ds[j]) PIME_CUDA(cudaMemcpyAsync(hs[j], ds[j], b, cudaMemcpyDeviceToHost, s));
    PIME_CUDA(cudaStreamSynchronize(s));
    return PIME_OK;
}

#ifdef PIME_PROFILE_WORKER
extern "C" int pime_debug_worker_prof(double *out16) {
    return cudaMemcpyFromSymbol(out16, pime::tc::g_worker_prof, 16 * sizeof(double)) == cudaSuccess ? 0 : -3;
}
#endif
