#include "rollout_ph.cuh"
using namespace pime;
extern "C" int pime_ph_rollout_f64(const pime_ph_config *cfg, const double *table, int64_t n, const pime_ph_state *st,
                                   const pime_rollout_args *args, void *stream) {
    return ph_rollout_impl<double>(cfg, table, n, st, args, stream);
}
