/*
 * pime_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  Not part of the product.
 *
 * A plain-C restatement of the reference's algorithm for the hot path (batched plant step +
 * P/PI prior + integrated-error observation + residual actor forward), following the numpy /
 * torch code of ruoqizzz/PIME line by line.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this file; the product library
 * (libpime_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself:
 * oracle/gen_golden.py imports the unmodified reference (via oracle/ref_loader.py) in the
 * dev container and commits its per-step tuples under tests/golden/; tests/test_oracle_golden.py
 * checks this file against them (and against the KATs in SURVEY.md section 8c).
 *
 * Build:  see oracle/Makefile  (gcc -O2 -ffp-contract=off -fPIC -shared; no fast-math, so every
 * double operation rounds exactly like the numpy float64 scalar ops it restates).
 *
 * All citations are file:line into the reference checkout.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PIME_REWARD_DISTANCE 0
#define PIME_REWARD_SQUARE 1
#define PIME_REWARD_SPARSE 2

/* ------------------------------------------------------------------------------------------
 * Configuration records (mirrors include/pime_b200.h field for field; kept separate on purpose
 * so that the oracle does not depend on any product header).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double A1, A2, G;          /* tank cross sections, gravity (gym_control/__init__.py:50-69)   */
    double sample_t;           /* 2.0                                                             */
    int32_t n_discrete;        /* 20 Euler sub-steps                                              */
    int32_t max_step;          /* 200                                                             */
    double P_max_action;       /* 10.0 (nonlinear_watertank.py:258-260)                           */
    int32_t reward_type;       /* distance / square_distance / sparse (:486-514)                  */
    int32_t has_integrator;    /* 1: ...UniformGoalIntegrator (:800-826), 0: base step (:274-297) */
    double z1;                 /* 1                                                               */
    double distance_threshold; /* 0.05                                                            */
    double integral_max;       /* 25.0 (:731-733)                                                 */
    double integral_punish;    /* 0.0                                                             */
} oracle_wt_cfg;

typedef struct {
    int32_t reward_type;
    int32_t integrator_mode; /* 0: no integrator obs (ph.py:479-485), 1: clipped (:320-348), 2: NoBound (:449-478) */
    int32_t max_episode_steps; /* gym TimeLimit, 50 (gym_control/__init__.py:6) */
    int32_t table_len;         /* len(MHCl) = 100000 */
    double act_low, act_high;  /* 0.0, 1.5 (ph.py:146-147) */
    double distance_threshold;
    double integral_max, integral_punish;
    double action_punishment, action_change_punishment; /* 0, 0 */
    double mhcl_step;          /* 1e-5 */
} oracle_ph_cfg;

static inline double clip_lo0(double v) {
    /* np.clip(v, -0., inf)  (observation_space.low = -ones*0, high = inf; nonlinear_watertank.py:735-742) */
    if (v < -0.0) v = -0.0;
    return v;
}

static inline double wt_reward(const oracle_wt_cfg *c, double h2, double r) {
    /* compute_reward (:486-514) with goal_distance (:64-72): np.linalg.norm of a scalar = sqrt(x*x) == |x| */
    double d = fabs(h2 - r);
    if (c->reward_type == PIME_REWARD_SPARSE) return -(double)(float)(d > c->distance_threshold ? 1.0f : 0.0f);
    if (c->reward_type == PIME_REWARD_DISTANCE) return -d * c->z1;
    return -(d * d) * c->z1; /* numpy float64 scalar ** 2 == d*d */
}

/* One env.step() of the water tank for n independent envs.
 * Restates NonLinearWaterTankUniformGoalIntegrator.step (nonlinear_watertank.py:800-826), the base
 * NonLinearWaterTank.step (:274-297) when has_integrator == 0, and the plant part of the Stacking
 * variant (:1122-1149).  noise1/noise2 are the two get_noise() draws (:810-811), may be NULL (= 0).
 * I may be NULL when has_integrator == 0. */
void pime_oracle_wt_step(const oracle_wt_cfg *c, int64_t n, double *h1, double *h2, const double *r, double *I,
                         int32_t *t, const double *a1, const double *a2, const double *Kp, const double *action,
                         const double *noise1, const double *noise2, double *reward, uint8_t *done) {
    const double delta_t = c->sample_t / (double)c->n_discrete; /* :151 */
    for (int64_t i = 0; i < n; ++i) {
        t[i] += 1;                                                               /* :801 */
        double u = action[i] * c->P_max_action / 2. + c->P_max_action / 2.;      /* :260 */
        double x1 = h1[i], x2 = h2[i];
        for (int k = 0; k < c->n_discrete; ++k) {                                /* :805-809 */
            double s1 = sqrt(2 * c->G * x1);
            double n1 = x1 + (-a1[i] / c->A1 * s1 + Kp[i] / c->A1 * u) * delta_t;
            double n2 = x2 + (a1[i] / c->A2 * s1 - a2[i] / c->A2 * sqrt(2 * c->G * x2)) * delta_t;
            x1 = clip_lo0(n1);
            x2 = clip_lo0(n2);
        }
        x1 += noise1 ? noise1[i] : 0.0;                                          /* :810-811 */
        x2 += noise2 ? noise2[i] : 0.0;
        x1 = clip_lo0(x1);                                                       /* :812-813 */
        x2 = clip_lo0(x2);
        double rew = wt_reward(c, x2, r[i]);                                     /* :815 */
        done[i] = (uint8_t)(t[i] >= c->max_step);                                /* :816-821 */
        if (c->has_integrator) {
            double integ = I[i] + (r[i] - x2);                                   /* :822-823 */
            rew += -c->integral_punish * fabs(integ);                            /* :824 */
            if (integ < -c->integral_max) integ = -c->integral_max;              /* :825 */
            if (integ > c->integral_max) integ = c->integral_max;
            I[i] = integ;
        }
        h1[i] = x1;
        h2[i] = x2;
        reward[i] = rew;
    }
}

/* --------------------------------------------------------------------------------------- pH */

/* PH1D.__init__ titration table (ph.py:72-84): 5 abs-Newton steps per grid point on the quartic in [H+],
 * warm-started from the previous grid point.  npow selects how H**k is evaluated: 0 = libm pow()
 * (what numpy float64 scalar ** int calls), 1 = repeated multiplication. */
void pime_oracle_ph_table(int32_t table_len, double mhcl_step, double kw, double kchem, double ka, double MNaOH,
                          double MHA, double MNH3, int32_t npow, double *pH) {
    double H = 1e-14 / MNaOH;
    for (int32_t i = 0; i < table_len; ++i) {
        double m = (double)i * mhcl_step; /* np.arange(0., 1, 1e-5)[i] */
        double ak = MNH3 - m + MNaOH + kchem + ka;
        double bk = (kchem + ka) * MNaOH - (kchem + ka) * m - kw + MNH3 * ka + kchem * ka - ka * MHA;
        double ck = MNaOH * kchem * ka - kw * (ka + kchem) - m * kchem * ka - ka * kchem * MHA;
        double dk = -kchem * ka * kw;
        for (int j = 0; j < 5; ++j) {
            double H2, H3, H4;
            if (npow == 0) {
                H2 = pow(H, 2.0); H3 = pow(H, 3.0); H4 = pow(H, 4.0);
            } else {
                H2 = H * H; H3 = H2 * H; H4 = H3 * H;
            }
            double num = H4 + ak * H3 + bk * H2 + ck * H + dk;
            double den = 4 * H3 + 3 * ak * H2 + 2 * bk * H + ck;
            H = fabs(H - num / den);
        }
        pH[i] = -1 * log10(H);
    }
}

/* PH1D.update_system (ph.py:114-121): ZOH discretisation of qc_V/(s+qww_V) at T = sample_t.
 * scipy realisation A=-qww_V, B=1, C=qc_V  =>  Ad = exp(-qww_V T), Bd = (1-Ad)/qww_V, Cd = qc_V. */
void pime_oracle_ph_update_system(int64_t n, double sample_t, const double *qww_V, const double *qc_V, double *A,
                                  double *B, double *C) {
    for (int64_t i = 0; i < n; ++i) {
        double a = exp(-qww_V[i] * sample_t);
        A[i] = a;
        B[i] = (1.0 - a) / qww_V[i];
        C[i] = qc_V[i];
    }
}

/* PH1D.observe_state (ph.py:187-189): first index whose MHCl entry is >= around(C*x, 5). Returns -1 when the
 * reference would raise IndexError (no such entry). */
int32_t pime_oracle_ph_index(const oracle_ph_cfg *c, double Cx) {
    double k = rint(Cx * 1e5);   /* np.around(v, 5) == rint(v*1e5)/1e5 */
    double target = k / 1e5;
    int64_t i = (int64_t)k;
    if (i < 0) i = 0;
    if (i > c->table_len) i = c->table_len;
    while (i > 0 && (double)(i - 1) * c->mhcl_step >= target) --i;
    while (i < c->table_len && (double)i * c->mhcl_step < target) ++i;
    return i >= c->table_len ? -1 : (int32_t)i;
}

static inline double ph_reward(const oracle_ph_cfg *c, double y, double r) {
    double d = fabs(y - r); /* ph.py:202-225 */
    if (c->reward_type == PIME_REWARD_SPARSE) return -(double)(float)(d > c->distance_threshold ? 1.0f : 0.0f);
    if (c->reward_type == PIME_REWARD_DISTANCE) return -d;
    return -(d * d);
}

/* PH1DUniformGoalIntegrator.step (ph.py:320-348), _NoBound.step (:449-478) and the TimeLimit wrapper.
 * Returns 0, or the (1-based) env index whose table lookup ran off the table (reference: IndexError). */
int64_t pime_oracle_ph_step(const oracle_ph_cfg *c, const double *table, int64_t n, double *x, double *y,
                            const double *r, double *I, int32_t *t, const double *A, const double *B, const double *C,
                            const double *action, double *reward, uint8_t *done) {
    for (int64_t i = 0; i < n; ++i) {
        double a = action[i];
        if (a < -1.0) a = -1.0;                                               /* :321 */
        if (a > 1.0) a = 1.0;
        t[i] += 1;                                                            /* :325 */
        double u = c->act_low + (c->act_high - c->act_low) * ((a - (-1.0)) / (1.0 - (-1.0))); /* :155-159 */
        double xn = A[i] * x[i] + B[i] * u;                                   /* :330 */
        int32_t k = pime_oracle_ph_index(c, C[i] * xn);                       /* :188 */
        if (k < 0) return i + 1;
        double yy = table[k];
        double rew = ph_reward(c, yy, r[i]);                                  /* :334 */
        rew -= c->action_punishment * fabs(u);                                /* :336 */
        /* action_change_punishment is 0 at every registered config; delta_u bookkeeping omitted (:322-324,337) */
        if (c->integrator_mode) {
            double integ = I[i] + (r[i] - yy);                                /* :339-340 */
            rew += -c->integral_punish * fabs(integ);                         /* :343 */
            if (c->integrator_mode == 1) {                                    /* :341 */
                if (integ < -c->integral_max) integ = -c->integral_max;
                if (integ > c->integral_max) integ = c->integral_max;
            }
            I[i] = integ;
        }
        x[i] = xn;
        y[i] = yy;
        reward[i] = rew;
        done[i] = (uint8_t)(t[i] >= c->max_episode_steps); /* gym TimeLimit */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ actor */

typedef struct {
    int32_t kind;      /* 0: ActorResidualPPO (net_residual.py:6-66), 1: ActorResidualIntegratorModularPPO (:138-205) */
    int32_t state_dim; /* S */
    int32_t mid_dim;   /* H */
    int32_t integrator_dim; /* modular only (1) */
} oracle_actor_cfg;

/* y[o] = b[o] + sum_i W[o*in+i] x[i]  (torch nn.Linear, weight [out,in], fp32) */
static void linear_f32(const float *W, const float *b, const float *x, int in, int out, float *y, int act_tanh) {
    for (int o = 0; o < out; ++o) {
        float acc = b[o];
        const float *w = W + (size_t)o * in;
        for (int i = 0; i < in; ++i) acc += w[i] * x[i];
        y[o] = act_tanh ? tanhf(acc) : acc;
    }
}

/* Plain actor parameter pack (fp32, torch layouts), in this order:
 *   W0[H,S] b0[H] W1[H,H] b1[H] W2[H,H] b2[H] W3[1,H] b3[1]
 * Modular actor parameter pack:
 *   Wo0[H,S-Di] bo0[H] Wo1[H/2,H] bo1[H/2] Wi0[H,Di] bi0[H] Wi1[H/2,H] bi1[H/2] Wn0[H,H] bn0[H] Wn1[1,H] bn1[1] */
int64_t pime_oracle_actor_param_count(const oracle_actor_cfg *c) {
    int64_t S = c->state_dim, H = c->mid_dim, D = c->integrator_dim;
    if (c->kind == 0) return H * S + H + 2 * (H * H + H) + H + 1;
    return H * (S - D) + H + (H / 2) * H + H / 2 + H * D + H + (H / 2) * H + H / 2 + H * H + H + H + 1;
}

/* a_avg = net(state) for one row (pre-tanh, pre-prior): net_residual.py:48-49 / :172-175 */
float pime_oracle_actor_avg(const oracle_actor_cfg *c, const float *p, const float *obs) {
    int S = c->state_dim, H = c->mid_dim;
    float buf0[1024], buf1[1024];
    float out;
    if (c->kind == 0) {
        const float *W0 = p, *b0 = W0 + H * S, *W1 = b0 + H, *b1 = W1 + H * H, *W2 = b1 + H, *b2 = W2 + H * H,
                    *W3 = b2 + H, *b3 = W3 + H;
        linear_f32(W0, b0, obs, S, H, buf0, 1);
        linear_f32(W1, b1, buf0, H, H, buf1, 1);
        linear_f32(W2, b2, buf1, H, H, buf0, 1);
        linear_f32(W3, b3, buf0, H, 1, &out, 0);
        return out;
    }
    int D = c->integrator_dim, So = S - D, Hh = H / 2;
    const float *Wo0 = p, *bo0 = Wo0 + H * So, *Wo1 = bo0 + H, *bo1 = Wo1 + Hh * H, *Wi0 = bo1 + Hh, *bi0 = Wi0 + H * D,
                *Wi1 = bi0 + H, *bi1 = Wi1 + Hh * H, *Wn0 = bi1 + Hh, *bn0 = Wn0 + H * H, *Wn1 = bn0 + H, *bn1 = Wn1 + H;
    float cat[1024];
    linear_f32(Wo0, bo0, obs, So, H, buf0, 1);          /* other_net[0..1]      (:151) */
    linear_f32(Wo1, bo1, buf0, H, Hh, cat, 1);          /* other_net[2..3]      (:152) */
    linear_f32(Wi0, bi0, obs + So, D, H, buf0, 1);      /* integrator_net[0..1] (:154) */
    linear_f32(Wi1, bi1, buf0, H, Hh, cat + Hh, 1);     /* integrator_net[2..3] (:155) */
    linear_f32(Wn0, bn0, cat, 2 * Hh, H, buf1, 1);      /* net[0..1]            (:157) */
    linear_f32(Wn1, bn1, buf1, H, 1, &out, 0);          /* net[2]               (:158) */
    return out;
}

void pime_oracle_actor_forward(const oracle_actor_cfg *c, const float *params, int64_t n, const float *obs /*[n,S]*/,
                               float *a_avg /*[n]*/) {
    for (int64_t i = 0; i < n; ++i) a_avg[i] = pime_oracle_actor_avg(c, params, obs + i * c->state_dim);
}

/* --------------------------------------------------------------------------------- rollouts */

/* Fused reference loop for the water tank: AgentResidualPPO.explore_env (agent_residual.py:52-69) when
 * deterministic == 0, get_episode_return (run.py:600-619) when deterministic == 1, run for T steps on n
 * independent envs that were already reset (state arrays hold the post-reset state).
 *
 *   obs32   = float32(obs)                                         elegantrl/env.py:46,72
 *   a_avg   = net(obs32)                    (fp32)                 net_residual.py:172-175
 *   a_raw   = a_avg + eps * exp(a_std_log)  (fp32)                 net_residual.py:176-180
 *   env_act = tanh_f32(a_raw) + obs32 @ priorK (fp64)              agent_residual.py:61
 *   (deterministic: env_act = float32(tanh(a_avg) + obs32@priorK32), net_residual.py:167-170)
 *
 * eps  [T,n] exploration noise (NULL = 0), pn1/pn2 [T,n] process noise (NULL = 0).
 * Replay rows (time-major): buf_state[T,n,S] float32, buf_other[T,n,4] = reward*scale, mask, a_raw, eps.
 * obs_mode: 0 = [h1,h2,r], 1 = [h1,h2,r,I], 2 = stacking (num_stack frames of [h1,h2,r], oldest first).
 * frames: [n, 3*num_stack] doubles for obs_mode 2 (in/out), else NULL.  params==NULL => a_avg = 0 (prior only). */
void pime_oracle_wt_rollout(const oracle_wt_cfg *c, const oracle_actor_cfg *ac, const float *params, float a_std_log,
                            const double *priorK, int32_t obs_mode, int32_t num_stack, int32_t deterministic, int64_t n,
                            int32_t T, double *h1, double *h2, const double *r, double *I, int32_t *t, const double *a1,
                            const double *a2, const double *Kp, double *frames, const float *eps, const double *pn1,
                            const double *pn2, double reward_scale, double gamma, float *buf_state, float *buf_other,
                            double *ep_return, double *env_action_out /*[T,n] or NULL*/) {
    const int S = obs_mode == 0 ? 3 : (obs_mode == 1 ? 4 : 3 * num_stack);
    const float a_std = expf(a_std_log);
    for (int64_t i = 0; i < n; ++i) {
        double Ii = I ? I[i] : 0.0;
        for (int32_t s = 0; s < T; ++s) {
            float obs32[64];
            if (obs_mode == 2) {
                for (int k = 0; k < S; ++k) obs32[k] = (float)frames[i * S + k];
            } else {
                obs32[0] = (float)h1[i]; obs32[1] = (float)h2[i]; obs32[2] = (float)r[i];
                if (obs_mode == 1) obs32[3] = (float)Ii;
            }
            float a_avg = params ? pime_oracle_actor_avg(ac, params, obs32) : 0.0f;
            float e = eps ? eps[(int64_t)s * n + i] : 0.0f;
            double env_act;
            float a_raw;
            if (deterministic) {
                float prior32 = 0.f;
                for (int k = 0; k < S; ++k) prior32 += obs32[k] * (float)priorK[k];
                a_raw = a_avg;
                env_act = (double)(tanhf(a_avg) + prior32);
            } else {
                double prior = 0.0;
                for (int k = 0; k < S; ++k) prior += (double)obs32[k] * priorK[k];
                a_raw = a_avg + e * a_std;
                env_act = (double)tanhf(a_raw) + prior;
            }
            double rew; uint8_t dn;
            double Itmp = Ii;
            pime_oracle_wt_step(c, 1, h1 + i, h2 + i, r + i, &Itmp, t + i, a1 + i, a2 + i, Kp + i, &env_act,
                                pn1 ? pn1 + (int64_t)s * n + i : NULL, pn2 ? pn2 + (int64_t)s * n + i : NULL, &rew, &dn);
            Ii = Itmp;
            if (obs_mode == 2) { /* frames.append([h1,h2,r]) (nonlinear_watertank.py:1145-1146) */
                memmove(frames + i * S, frames + i * S + 3, sizeof(double) * (S - 3));
                frames[i * S + S - 3] = h1[i]; frames[i * S + S - 2] = h2[i]; frames[i * S + S - 1] = r[i];
            }
            if (buf_state) {
                float *bs = buf_state + ((int64_t)s * n + i) * S;
                for (int k = 0; k < S; ++k) bs[k] = obs32[k];
                float *bo = buf_other + ((int64_t)s * n + i) * 4;
                bo[0] = (float)(rew * reward_scale); bo[1] = dn ? 0.0f : (float)gamma; bo[2] = a_raw; bo[3] = e;
            }
            if (env_action_out) env_action_out[(int64_t)s * n + i] = env_act;
            if (ep_return) ep_return[i] += rew;
        }
        if (I) I[i] = Ii;
    }
}

/* Same loop for the pH plant (obs = [y, r, I] / [y, r]); table lookups as pime_oracle_ph_step. */
int64_t pime_oracle_ph_rollout(const oracle_ph_cfg *c, const double *table, const oracle_actor_cfg *ac, const float *params,
                               float a_std_log, const double *priorK, int32_t deterministic, int64_t n, int32_t T, double *x,
                               double *y, const double *r, double *I, int32_t *t, const double *A, const double *B,
                               const double *C, const float *eps, double reward_scale, double gamma, float *buf_state,
                               float *buf_other, double *ep_return, double *env_action_out) {
    const int S = c->integrator_mode ? 3 : 2;
    const float a_std = expf(a_std_log);
    for (int64_t i = 0; i < n; ++i) {
        double Ii = I ? I[i] : 0.0;
        for (int32_t s = 0; s < T; ++s) {
            float obs32[3];
            obs32[0] = (float)y[i]; obs32[1] = (float)r[i];
            if (S == 3) obs32[2] = (float)Ii;
            float a_avg = params ? pime_oracle_actor_avg(ac, params, obs32) : 0.0f;
            float e = eps ? eps[(int64_t)s * n + i] : 0.0f;
            double env_act; float a_raw;
            if (deterministic) {
                float prior32 = 0.f;
                for (int k = 0; k < S; ++k) prior32 += obs32[k] * (float)priorK[k];
                a_raw = a_avg;
                env_act = (double)(tanhf(a_avg) + prior32);
            } else {
                double prior = 0.0;
                for (int k = 0; k < S; ++k) prior += (double)obs32[k] * priorK[k];
                a_raw = a_avg + e * a_std;
                env_act = (double)tanhf(a_raw) + prior;
            }
            double rew; uint8_t dn; double Itmp = Ii;
            int64_t err = pime_oracle_ph_step(c, table, 1, x + i, y + i, r + i, &Itmp, t + i, A + i, B + i, C + i, &env_act,
                                              &rew, &dn);
            if (err) return i + 1;
            Ii = Itmp;
            if (buf_state) {
                float *bs = buf_state + ((int64_t)s * n + i) * S;
                for (int k = 0; k < S; ++k) bs[k] = obs32[k];
                float *bo = buf_other + ((int64_t)s * n + i) * 4;
                bo[0] = (float)(rew * reward_scale); bo[1] = dn ? 0.0f : (float)gamma; bo[2] = a_raw; bo[3] = e;
            }
            if (env_action_out) env_action_out[(int64_t)s * n + i] = env_act;
            if (ep_return) ep_return[i] += rew;
        }
        if (I) I[i] = Ii;
    }
    return 0;
}

/* -------------------------------------------------------------------------- Philox4x32-10 */
/* Counter-based RNG used by the product's reset / noise kernels (a new design: the reference draws from
 * numpy's global Mersenne Twister, nonlinear_watertank.py:272,890-894,912-913, which cannot be matched on a
 * device).  Restated here so the integer stream and the uniform->parameter mapping are checked bit-exactly. */
static inline void philox_round(uint32_t ctr[4], const uint32_t key[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ key[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ key[1];
    uint32_t n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
}

void pime_oracle_philox4x32(uint64_t seed, uint64_t index, uint32_t tick, uint32_t stream, uint32_t out[4]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), tick, stream};
    for (int i = 0; i < 10; ++i) {
        philox_round(ctr, key);
        key[0] += 0x9E3779B9u;
        key[1] += 0xBB67AE85u;
    }
    memcpy(out, ctr, sizeof(uint32_t) * 4);
}

static inline double u01_53(uint32_t lo, uint32_t hi) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

/* Reset draws for env `index`, episode `episode`: returns 6 uniforms in [0,1) in the order
 * a1, a2, Kp, h1, h2, r  (water tank: sample_parameters :890-894 then reset_all :902-916)  or
 * qww_V, qc_V, x, r, -, -  (pH: sample_parameters :409-410 then reset_all :412-426). */
void pime_oracle_reset_uniforms(uint64_t seed, uint64_t index, uint32_t episode, double u[6]) {
    uint32_t w[4];
    for (uint32_t s = 0; s < 3; ++s) {
        pime_oracle_philox4x32(seed, index, episode, s, w); /* streams 0,1,2 = reset */
        u[2 * s] = u01_53(w[0], w[1]);
        u[2 * s + 1] = u01_53(w[2], w[3]);
    }
}
