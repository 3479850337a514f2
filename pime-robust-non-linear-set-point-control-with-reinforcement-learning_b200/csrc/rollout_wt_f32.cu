#include "rollout_wt.cuh"
#include "host_pipe.cuh"
using namespace pime;
extern "C" int pime_wt_rollout_f32(const pime_wt_config *cfg, int64_t n, const pime_wt_state *st, const pime_rollout_args *args,
                                   void *stream) {
    return wt_rollout_impl<float>(cfg, n, st, args, stream);
}

// Host-buffer entry: H2D of the per-env state, fused rollout, D2H of ep_return (+ final state), pipelined over env slices
// (host_pipe.cuh).  The stacking observation keeps its frames [3 num_stack][n] with stride n: one slice.
extern "C" int pime_wt_rollout_host_f32(const pime_wt_config *cfg, int64_t n, const pime_wt_state *h, const pime_wt_state *d,
                                        const pime_rollout_args *args, float *ep_return_host, void *stream) {
    PIME_REQUIRE(cfg && h && d && args, "null pointer");
    PIME_REQUIRE(d->ep_return, "device ep_return scratch is required");
    PIME_REQUIRE(n >= 0, "negative n");
    if (int rc = require_device()) return rc;
    const HostArr arr[9] = {{h->h1, d->h1, 4, true}, {h->h2, d->h2, 4, true}, {h->r, d->r, 4, true}, {h->I, d->I, 4, true},
                            {h->a1, d->a1, 4, false}, {h->a2, d->a2, 4, false}, {h->Kp, d->Kp, 4, false},
                            {h->t, d->t, 4, false}, {h->episode, d->episode, 4, false}};
    const bool single = cfg->obs_mode == PIME_WT_OBS_STACKING;
    auto launch = [&](int64_t off, int64_t cnt, const pime_rollout_args &a) {
        pime_wt_state st = *d;
        auto sh = [&](void *p, int elem) { return p ? (void *)((char *)p + off * elem) : nullptr; };
        st.h1 = sh(d->h1, 4); st.h2 = sh(d->h2, 4); st.r = sh(d->r, 4); st.I = sh(d->I, 4);
        st.a1 = sh(d->a1, 4); st.a2 = sh(d->a2, 4); st.Kp = sh(d->Kp, 4);
        st.t = (int32_t *)sh(d->t, 4); st.episode = (uint32_t *)sh(d->episode, 4);
        st.ep_return = sh(d->ep_return, 4); st.last_h1 = sh(d->last_h1, 4); st.last_h2 = sh(d->last_h2, 4);
        pime_rollout_args b = a;
        shift_step_buffers(b, off, cfg->obs_mode == PIME_WT_OBS_INTEGRATOR ? 4 : 3, 4);   // (stacking: one slice, off = 0)
        return wt_rollout_impl<float>(cfg, cnt, &st, &b, stream);
    };
    return host_pipelined_rollout(n, arr, 9, (float *)d->ep_return, ep_return_host, args, (cudaStream_t)stream, launch, single);
}

// the slicing the host-buffer entries would use for n envs on the current device (148 SMs assumed when there is none):
// out2 = {number of slices, envs per slice}.  Host arithmetic only.
extern "C" int pime_host_slice_plan(int64_t n, int32_t single, int64_t *out2) {
    PIME_REQUIRE(out2 && n >= 0, "null output / negative n");
    const int slices = n == 0 ? 1 : host_slice_count(n, single != 0);
    out2[0] = slices;
    out2[1] = host_slice_len(n, slices);
    return PIME_OK;
}

// tests / tuning: force the number of env slices of the host-buffer entries (0 = automatic)
extern "C" int pime_set_host_slices(int32_t slices) {
    PIME_REQUIRE(slices >= 0 && slices <= kMaxSlices, "0 <= slices <= 8");
    host_slices_override() = slices;
    return PIME_OK;
}

#ifdef PIME_PROFILE_WORKER
extern "C" int pime_debug_worker_prof(double *out16) {
    return cudaMemcpyFromSymbol(out16, pime::tc::g_worker_prof, 16 * sizeof(double)) == cudaSuccess ? 0 : -3;
}
#endif
