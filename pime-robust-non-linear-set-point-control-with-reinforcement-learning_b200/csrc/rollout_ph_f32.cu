#include "rollout_ph.cuh"
using namespace pime;
extern "C" int pime_ph_rollout_f32(const pime_ph_config *cfg, const float *table, int64_t n, const pime_ph_state *st,
                                   const pime_rollout_args *args, void *stream) {
    return ph_rollout_impl<float>(cfg, table, n, st, args, stream);
}

// Host-buffer entry (configs[3] end to end): H2D of the per-env state, fused rollout, D2H of ep_return + the final state.
// x, A, B are double arrays in the float flavour (include/pime_b200.h), everything else float / int32.
extern "C" int pime_ph_rollout_host_f32(const pime_ph_config *cfg, const float *table, int64_t n, const pime_ph_state *h,
                                        const pime_ph_state *d, const pime_rollout_args *args, float *ep_return_host, void *stream) {
    PIME_REQUIRE(cfg && table && h && d && args, "null pointer");
    PIME_REQUIRE(d->ep_return, "device ep_return scratch is required");
    if (int rc = require_device()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t b4 = (size_t)n * 4, b8 = (size_t)n * 8;
    struct { void *hp, *dp; size_t bytes; bool back; } arr[] = {
        {h->x, d->x, b8, true}, {h->y, d->y, b4, true}, {h->r, d->r, b4, true}, {h->I, d->I, b4, true},
        {h->A, d->A, b8, false}, {h->B, d->B, b8, false}, {h->C, d->C, b4, false},
        {h->qww_V, d->qww_V, b4, false}, {h->qc_V, d->qc_V, b4, false},
        {h->t, d->t, b4, false}, {h->episode, d->episode, b4, false}};
    for (auto &a : arr)
        if (a.hp && a.dp) PIME_CUDA(cudaMemcpyAsync(a.dp, a.hp, a.bytes, cudaMemcpyHostToDevice, s));
    PIME_CUDA(cudaMemsetAsync(d->ep_return, 0, b4, s));
    if (int rc = ph_rollout_impl<float>(cfg, table, n, d, args, stream)) return rc;
    if (ep_return_host) PIME_CUDA(cudaMemcpyAsync(ep_return_host, d->ep_return, b4, cudaMemcpyDeviceToHost, s));
    for (auto &a : arr)
        if (a.back && a.hp && a.dp) PIME_CUDA(cudaMemcpyAsync(a.hp, a.dp, a.bytes, cudaMemcpyDeviceToHost, s));
    PIME_CUDA(cudaStreamSynchronize(s));
    return PIME_OK;
}
