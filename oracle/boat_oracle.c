/*
 * boat_oracle.c -- CPU restatement (plain C, fp64) of the reference's environment
 * step hot path.  TEST INFRASTRUCTURE ONLY (see boat_oracle.h): the product path
 * never links or calls this file.
 *
 * Parity status: PINNED against the reference's recorded fixtures and against the
 * live reference classes (see boat_oracle.h).
 *
 * The arithmetic follows the reference's operation order (Python evaluates a*b*c
 * as (a*b)*c) so that results agree with the numpy-scalar original to rounding
 * noise; build with -ffp-contract=off so gcc does not fuse multiplies and adds.
 */
#include "boat_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* control_blocks.py:5-36  Integrator                                        */
/* ------------------------------------------------------------------------- */
typedef struct {
    double dt, initial_value, lower, upper, last_stored;
    long counter;
} integrator;

static void integrator_init(integrator *g, double initial_value, double lower, double upper,
                            double dt)
{
    g->dt = dt;                 /* control_blocks.py:7 (0.1) unless overwritten  */
    g->initial_value = initial_value;
    g->lower = lower;
    g->upper = upper;
    g->last_stored = initial_value; /* output_signal = [initial_value]  :14     */
    g->counter = 0;
}

/* control_blocks.py:16-36: call 0 returns initial_value and ignores the input; the
 * STORED value is clamped, the RETURNED value is not. */
static double integrate_signal(integrator *g, double x)
{
    double y;
    if (g->counter == 0)
        y = g->initial_value;
    else
        y = x * g->dt + g->last_stored;
    if (y <= g->lower)
        g->last_stored = g->lower;
    else if (y >= g->upper)
        g->last_stored = g->upper;
    else
        g->last_stored = y;
    g->counter += 1;
    return y;
}

static double sign_d(double x) { return (double)((x > 0.0) - (x < 0.0)); } /* np.sign */
static double square_d(double x) { return x * x; }                         /* np.square */

/* ------------------------------------------------------------------------- */
/* wind.py                                                                   */
/* ------------------------------------------------------------------------- */
int oracle_wind_length(const oracle_params *p)
{
    return (int)(p->t_max / p->dt); /* wind.py:14-15 */
}

/* Natural restatement of scipy.interpolate.interp1d(kind='cubic') as called at
 * wind.py:82-84: scipy builds make_interp_spline(x, y, k=3) whose default
 * boundary condition is not-a-knot, i.e. the interpolating C2 cubic whose third
 * derivative is continuous across x[1] and x[n-2].  (scipy 1.9.3 is pinned in
 * requirements.txt:29 and is not vendored; the published algorithm is restated
 * here through the second-derivative ("moment") form and pinned by the
 * reference's wind.csv fixtures, which it reproduces to ~1e-15.)
 *
 * Unknowns M[i] = S''(x[i]).  Interior rows: h/6 M[i-1] + 2h/3 M[i] + h/6 M[i+1]
 * = (y[i+1]-2y[i]+y[i-1])/h (uniform h).  Not-a-knot rows: M[0]-2M[1]+M[2] = 0 and
 * M[n-3]-2M[n-2]+M[n-1] = 0.  Solved by dense Gaussian elimination with partial
 * pivoting (n is 8). */
static int spline_moments(const double *y, int n, double h, double *M)
{
    double *A = (double *)calloc((size_t)n * (n + 1), sizeof(double));
    if (!A) return -1;
#define AT(r, c) A[(r) * (n + 1) + (c)]
    AT(0, 0) = 1.0; AT(0, 1) = -2.0; AT(0, 2) = 1.0; AT(0, n) = 0.0;
    for (int i = 1; i < n - 1; ++i) {
        AT(i, i - 1) = h / 6.0;
        AT(i, i) = 2.0 * h / 3.0;
        AT(i, i + 1) = h / 6.0;
        AT(i, n) = (y[i + 1] - 2.0 * y[i] + y[i - 1]) / h;
    }
    AT(n - 1, n - 3) = 1.0; AT(n - 1, n - 2) = -2.0; AT(n - 1, n - 1) = 1.0; AT(n - 1, n) = 0.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(AT(r, c)) > fabs(AT(piv, c))) piv = r;
        if (piv != c)
            for (int k = 0; k <= n; ++k) { double t = AT(c, k); AT(c, k) = AT(piv, k); AT(piv, k) = t; }
        for (int r = c + 1; r < n; ++r) {
            double f = AT(r, c) / AT(c, c);
            for (int k = c; k <= n; ++k) AT(r, k) -= f * AT(c, k);
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double s = AT(r, n);
        for (int k = r + 1; k < n; ++k) s -= AT(r, k) * M[k];
        M[r] = s / AT(r, r);
    }
#undef AT
    free(A);
    return 0;
}

int oracle_random_curve(const double *knots, int n_knots, int L, double *out)
{
    if (n_knots < 4) return -1; /* wind.py:73-75 ValueError */
    /* fixed_points = np.linspace(0, L, num=n_knots)            wind.py:76-77 */
    const double h = (double)L / (double)(n_knots - 1);
    double *M = (double *)malloc(sizeof(double) * (size_t)n_knots);
    if (!M || spline_moments(knots, n_knots, h, M) != 0) { free(M); return -1; }
    /* complete_range = np.linspace(0, L, num=L, endpoint=True) wind.py:80-81:
     * numpy computes i*step with step = L/(L-1) and pins the last sample to L. */
    const double step = (double)L / (double)(L - 1);
    int any_outside = 0;
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < L; ++i) {
        double x = (i == L - 1) ? (double)L : (double)i * step;
        int j = (int)floor(x / h);
        if (j > n_knots - 2) j = n_knots - 2;
        if (j < 0) j = 0;
        double t = x - (double)j * h;
        double b = (knots[j + 1] - knots[j]) / h - h * (2.0 * M[j] + M[j + 1]) / 6.0;
        double c = M[j] / 2.0;
        double d = (M[j + 1] - M[j]) / (6.0 * h);
        double v = knots[j] + t * (b + t * (c + t * d));
        out[i] = v;
        if (v < 0.0 || v > 1.0) any_outside = 1; /* wind.py:87 */
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    if (any_outside) /* wind.py:88-89 */
        for (int i = 0; i < L; ++i) out[i] = (out[i] - mn) / (mx - mn);
    free(M);
    return 0;
}

int oracle_generate_wind(const oracle_params *p, const double *knots_a, const double *knots_b,
                         double *wv, double *wa)
{
    const int L = oracle_wind_length(p);
    int rc = 0;
    switch (p->experiment) {
    case 1: /* wind.py:31-33 */
    case 2: /* wind.py:35-37 */
        for (int i = 0; i < L; ++i) { wv[i] = 0.0; wa[i] = 0.0; }
        break;
    case 3: { /* wind.py:40-45 */
        double angle = p->direction * (M_PI / 180.0);
        for (int i = 0; i < L; ++i) { wv[i] = p->max_velocity; wa[i] = angle; }
        break;
    }
    case 4: { /* wind.py:47-51 */
        rc = oracle_random_curve(knots_a, p->fixed_points, L, wv);
        if (rc) return rc;
        double angle = p->direction * (M_PI / 180.0);
        for (int i = 0; i < L; ++i) { wv[i] = wv[i] * p->max_velocity; wa[i] = angle; }
        break;
    }
    case 5: { /* wind.py:53-58, rect_random_curve :92-99 with middle = 0.5 */
        rc = oracle_random_curve(knots_a, p->fixed_points, L, wa);
        if (rc) return rc;
        const double middle = 0.5;
        for (int i = 0; i < L; ++i) {
            double r = (wa[i] <= middle / 2.0) ? 0.0 : 1.0;
            wa[i] = (r * M_PI) + M_PI / 2.0;
            wv[i] = p->max_velocity;
        }
        break;
    }
    case 6: { /* wind.py:60-63: velocity curve is drawn first */
        rc = oracle_random_curve(knots_a, p->fixed_points, L, wv);
        if (rc) return rc;
        rc = oracle_random_curve(knots_b, p->fixed_points, L, wa);
        if (rc) return rc;
        for (int i = 0; i < L; ++i) { wv[i] = wv[i] * p->max_velocity; wa[i] = wa[i] * M_PI * 2.0; }
        break;
    }
    default:
        return -2; /* wind.py:65-67 ValueError */
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* boat_env.py  Boat + BoatEnv                                               */
/* ------------------------------------------------------------------------- */
struct oracle_env {
    oracle_params p;
    int L;
    double *wind_velocity, *wind_angle;
    /* Boat.__init__ boat_env.py:144-201 */
    int s_y_start;
    double t;
    long index;
    integrator a_x_i, v_x_i, a_y_i, v_y_i, a_r_i, v_r_i;
    double n, rudder_angle, fuel;
    double a_x, v_x, s_x, a_y, v_y, s_y, a_r, v_r, s_r, v, drift_angle;
    double out_of_bounds;
    /* BoatEnv */
    double action, reward, episode_reward;
};

oracle_env *oracle_env_create(const oracle_params *p)
{
    oracle_env *e = (oracle_env *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    e->p = *p;
    e->L = oracle_wind_length(p);
    e->wind_velocity = (double *)calloc((size_t)e->L, sizeof(double));
    e->wind_angle = (double *)calloc((size_t)e->L, sizeof(double));
    if (!e->wind_velocity || !e->wind_angle) { oracle_env_destroy(e); return NULL; }
    return e;
}

void oracle_env_destroy(oracle_env *e)
{
    if (!e) return;
    free(e->wind_velocity);
    free(e->wind_angle);
    free(e);
}

/* boat_env.py:283-306 */
static void get_kinematics(oracle_env *e)
{
    e->v = sqrt(square_d(e->v_x) + square_d(e->v_y));
    e->drift_angle = atan2(e->v_x, e->v_y); /* x first, boat_env.py:289 */
    /* turning_rate (:292-294) is computed by the reference but never read. */
    e->s_r = integrate_signal(&e->v_r_i, e->v_r);
    double direction = e->drift_angle - e->s_r;
    double v_x_new = sin(direction) * e->v;
    e->s_x = integrate_signal(&e->v_x_i, v_x_new);
    double v_y_new = cos(direction) * e->v;
    e->s_y = integrate_signal(&e->v_y_i, v_y_new);
}

/* boat_env.py:308-326 */
static double normalize(double value, double lo, double hi) { return (value - lo) / (hi - lo); }

static void return_state(const oracle_env *e, double *s)
{
    const oracle_params *p = &e->p;
    s[0] = normalize(e->s_x, 0.0, p->goal_line);
    s[1] = normalize(e->v_x, 0.0, 5.0);
    s[2] = normalize(e->a_x, 0.0, 0.025);
    s[3] = normalize(e->s_y, -p->track_width, p->track_width);
    s[4] = normalize(e->v_y, 0.0, 2.0);
    s[5] = normalize(e->a_y, 0.0, 0.37);
    s[6] = normalize(e->s_r, 0.0, 2.0 * M_PI);
    s[7] = normalize(e->v_r, 0.0, 8.5e-3);
    s[8] = normalize(e->a_r, 0.0, 1.4e-5);
    s[9] = normalize(e->rudder_angle, -M_PI / 3.0, M_PI / 3.0);
    s[10] = normalize(e->fuel, 0.0, p->fuel);
}

static void boat_init(oracle_env *e, int s_y_start)
{
    const oracle_params *p = &e->p;
    e->s_y_start = s_y_start; /* boat_env.py:147-150 (always drawn) */
    e->t = 0.0;
    e->index = 0;
    const double inf = INFINITY;
    integrator_init(&e->a_x_i, 3.0, -inf, inf, p->dt); /* :158-159 */
    integrator_init(&e->v_x_i, 0.0, -inf, inf, p->dt); /* :160-161 */
    integrator_init(&e->a_y_i, 0.0, -inf, inf, p->dt); /* :163-164 */
    integrator_init(&e->v_y_i, p->experiment == 2 ? (double)s_y_start : 0.0, -inf, inf,
                    p->dt);                            /* :166-170 */
    integrator_init(&e->a_r_i, 0.0, -inf, inf, p->dt); /* :172-173 */
    integrator_init(&e->v_r_i, 0.0, -inf, inf, p->dt); /* :174-175 */
    e->n = 20.0;          /* :178 */
    e->rudder_angle = 0.0;
    e->fuel = p->fuel;    /* :180 */
    e->a_x = e->v_x = e->s_x = 0.0;
    e->a_y = e->v_y = e->s_y = 0.0;
    e->a_r = e->v_r = e->s_r = 0.0;
    e->v = 0.0;
    e->drift_angle = 0.0;
    get_kinematics(e);    /* :198 -- burns call 0 of the three position integrators */
    e->out_of_bounds = p->track_width + p->oob_offset; /* :200-201 */
    e->action = 0.0;
    e->reward = 0.0;
}

int oracle_env_reset(oracle_env *e, int s_y_start, const double *knots_a, const double *knots_b,
                     double *obs)
{
    int rc = oracle_generate_wind(&e->p, knots_a, knots_b, e->wind_velocity, e->wind_angle);
    if (rc) return rc;
    boat_init(e, s_y_start);
    e->episode_reward = 0.0; /* boat_env.py:122 */
    if (obs) return_state(e, obs);
    return 0;
}

int oracle_env_reset_tables(oracle_env *e, int s_y_start, const double *wv, const double *wa,
                            double *obs)
{
    memcpy(e->wind_velocity, wv, sizeof(double) * (size_t)e->L);
    memcpy(e->wind_angle, wa, sizeof(double) * (size_t)e->L);
    boat_init(e, s_y_start);
    e->episode_reward = 0.0;
    if (obs) return_state(e, obs);
    return 0;
}

/* boat_env.py:213-239 */
static void eom_longitudinal(oracle_env *e)
{
    const oracle_params *p = &e->p;
    double F_R = square_d(e->v_x) * p->c_r_front * 0.5 * p->rho * p->boat_area_front;
    double v_x_w = e->v_x * (1.0 - p->wake_friction);
    double J = 0.0;
    if (e->n != 0.0) J = v_x_w / (e->n * p->propeller_diameter);
    double KT = sin(J);
    double F_T = KT * square_d(e->n) * p->rho * pow(p->propeller_diameter, 4.0) *
                 (1.0 - p->thrust_deduction);
    double F_C = e->v_y * (p->boat_m + p->boat_m_y) * e->v_r;
    double w = e->wind_velocity[e->index];
    double F_W_unangled =
        square_d(w) * sign_d(w) * p->c_r_front * 0.5 * p->rho * p->boat_area_front;
    double F_W = F_W_unangled * cos(e->wind_angle[e->index]);
    e->a_x = (-F_R + F_T + F_C + F_W) / (p->boat_m + p->boat_m_x);
}

/* boat_env.py:241-265 */
static void eom_transverse(oracle_env *e)
{
    const oracle_params *p = &e->p;
    double v_y_sign = sign_d(e->v_y);
    double F_R = square_d(e->v_y) * p->c_r_side * 0.5 * p->rho * p->boat_area_side * v_y_sign;
    double F_RU = square_d(e->v_x) * p->c_r_front * 0.5 * p->rho * p->rudder_area;
    F_RU = sin(e->rudder_angle) * F_RU;
    double F_C = e->v_x * (p->boat_m + p->boat_m_x) * e->v_r;
    double w = e->wind_velocity[e->index];
    double F_W_unangled =
        square_d(w) * sign_d(w) * p->c_r_side * 0.5 * p->rho * p->boat_area_side;
    double F_W = F_W_unangled * sin(e->wind_angle[e->index]);
    e->a_y = (-F_R + F_RU + F_C + F_W) / (p->boat_m + p->boat_m_y);
}

/* boat_env.py:267-281 */
static void eom_yawning(oracle_env *e)
{
    const oracle_params *p = &e->p;
    double v_r_sign = sign_d(e->v_r);
    double v_x_sign = sign_d(e->v_x);
    double M_hull = square_d(e->v_r) * p->c_r_side * 0.5 * p->rho * p->boat_area_side *
                    p->boat_l * 5.0 * v_r_sign;
    double M_rudder = square_d(e->v_x) * p->c_r_side * 0.5 * p->rho * p->rudder_area *
                      sin(e->rudder_angle) * (p->boat_b / 2.0) * v_x_sign;
    e->a_r = (-M_hull + M_rudder) / (p->boat_I + p->boat_Iz);
}

/* reward_functions.py:42-57 with the constructor arguments of boat_env.py:16-22
 * (y_a = 0.03, y_b = 3.4); f_x == 0 in the current revision (:48-49). */
static double exponential_reward(const oracle_params *p, double y)
{
    const double y_a = 0.03, y_b = 3.4;
    double f_y = (fabs(y) / p->track_width) /
                 (1.0 + exp((-y_a / y_b) * (fabs(y) - (p->track_width * 0.2))));
    return 0.0 - f_y;
}

int oracle_env_step(oracle_env *e, double action, double *obs, double *reward, int *done,
                    int *term_code)
{
    const oracle_params *p = &e->p;
    if (e->index >= e->L) return -3; /* IndexError in wind.get_wind */
    e->action = action;              /* boat_env.py:68 */
    e->t += p->dt;                   /* :69 */
    e->fuel -= 1.0;                  /* :70 */
    if (p->test_mode == 0) e->rudder_angle += action / 10.0; /* :72-73 */

    /* run_model_step :203-211 */
    eom_longitudinal(e);
    e->v_x = integrate_signal(&e->a_x_i, e->a_x);
    eom_transverse(e);
    e->v_y = integrate_signal(&e->a_y_i, e->a_y);
    eom_yawning(e);
    e->v_r = integrate_signal(&e->a_r_i, e->a_r);
    get_kinematics(e);
    e->index += 1;

    if (obs) return_state(e, obs); /* :77 */
    double r = exponential_reward(p, e->s_y); /* :80-81 */

    int d = 0, code = ORACLE_TERM_NONE; /* :84-105 */
    if (e->s_x >= p->goal_line) { d = 1; code = ORACLE_TERM_GOAL; r += 1000.0; }
    else if (fabs(e->s_y) > e->out_of_bounds || e->s_x < 0.0) { d = 1; code = ORACLE_TERM_OOB; }
    else if (e->fuel < 0.0) { d = 1; code = ORACLE_TERM_FUEL; }
    else if (p->t_max <= e->t) { d = 1; code = ORACLE_TERM_TIMEOUT; }
    else if (e->rudder_angle > M_PI / 3.0 || e->rudder_angle < -M_PI / 3.0) {
        d = 1; code = ORACLE_TERM_RUDDER;
    }
    if (e->rudder_angle > M_PI / 4.0 || e->rudder_angle < -M_PI / 4.0) /* :107-108 */
        r -= fabs(e->rudder_angle) * 100.0;
    if (fabs(e->s_r) > M_PI / 2.0) r -= 1.0; /* :110-111 */
    e->reward = r;
    e->episode_reward += r; /* :113 */
    if (reward) *reward = r;
    if (done) *done = d;
    if (term_code) *term_code = code;
    return 0;
}

void oracle_env_all_data(const oracle_env *e, double *o)
{
    o[0] = e->s_x; o[1] = e->s_y; o[2] = e->v_x; o[3] = e->v_y; o[4] = e->s_r;
    o[5] = e->action; o[6] = e->reward; o[7] = e->rudder_angle;
}
const double *oracle_env_wind_velocity(const oracle_env *e) { return e->wind_velocity; }
const double *oracle_env_wind_angle(const oracle_env *e) { return e->wind_angle; }
double oracle_env_episode_reward(const oracle_env *e) { return e->episode_reward; }

typedef struct {
    const oracle_params *p;
    int n_envs, n_steps, n_episodes, auto_reset, tid, n_threads;
    const double *actions;
    const int *s_y_start;
    const double *knots;
    double *obs_out, *reward_out;
    unsigned char *done_out, *term_out;
    int *ep_len_out;
    double *final_state_out;
    long long total;
    int err;
} rollout_job;

static void *rollout_worker(void *arg)
{
    rollout_job *jb = (rollout_job *)arg;
    const oracle_params *p = jb->p;
    const int fp = p->fixed_points, n_envs = jb->n_envs;
    oracle_env *e = oracle_env_create(p);
    if (!e) { jb->err = -10; return NULL; }
    /* envs are dealt to threads in blocks of 16 (round-robin) */
    for (int base = jb->tid * 16; base < n_envs; base += jb->n_threads * 16) {
        for (int i = base; i < base + 16 && i < n_envs; ++i) {
            int ep = 0, first_len = -1, rc, steps_in_ep = 0;
            const double *k = jb->knots + ((size_t)ep * n_envs + i) * 2 * fp;
            rc = oracle_env_reset(e, jb->s_y_start[(size_t)ep * n_envs + i], k, k + fp, NULL);
            if (rc) { jb->err = rc; continue; }
            for (int s = 0; s < jb->n_steps; ++s) {
                double obs[11], r;
                int d, code;
                rc = oracle_env_step(e, jb->actions[(size_t)s * n_envs + i], obs, &r, &d, &code);
                if (rc) { jb->err = rc; break; }
                ++steps_in_ep;
                ++jb->total;
                size_t o = (size_t)s * n_envs + i;
                if (jb->obs_out) memcpy(jb->obs_out + o * 11, obs, sizeof(obs));
                if (jb->reward_out) jb->reward_out[o] = r;
                if (jb->done_out) jb->done_out[o] = (unsigned char)d;
                if (jb->term_out) jb->term_out[o] = (unsigned char)code;
                if (d) {
                    if (first_len < 0) first_len = steps_in_ep;
                    if (jb->auto_reset) {
                        ++ep;
                        if (ep >= jb->n_episodes) { jb->err = -11; break; }
                        k = jb->knots + ((size_t)ep * n_envs + i) * 2 * fp;
                        rc = oracle_env_reset(e, jb->s_y_start[(size_t)ep * n_envs + i], k, k + fp,
                                              NULL);
                        if (rc) { jb->err = rc; break; }
                        steps_in_ep = 0;
                    }
                }
            }
            if (jb->ep_len_out) jb->ep_len_out[i] = first_len;
            if (jb->final_state_out) {
                double *f = jb->final_state_out + (size_t)i * 8;
                f[0] = e->v_x; f[1] = e->v_y; f[2] = e->v_r; f[3] = e->rudder_angle;
                f[4] = e->s_x; f[5] = e->s_y; f[6] = e->s_r; f[7] = e->episode_reward;
            }
        }
    }
    oracle_env_destroy(e);
    return NULL;
}

long long oracle_rollout(const oracle_params *p, int n_envs, int n_steps, int n_episodes,
                         int auto_reset, const double *actions, const int *s_y_start,
                         const double *knots, double *obs_out, double *reward_out,
                         unsigned char *done_out, unsigned char *term_out, int *ep_len_out,
                         double *final_state_out, int n_threads)
{
    if (n_threads <= 0) {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        n_threads = nc > 0 ? (int)nc : 1;
    }
    if (n_threads > (n_envs + 15) / 16) n_threads = (n_envs + 15) / 16;
    if (n_threads < 1) n_threads = 1;
    rollout_job *jobs = (rollout_job *)calloc((size_t)n_threads, sizeof(rollout_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -10; }
    for (int t = 0; t < n_threads; ++t) {
        rollout_job jb = { p, n_envs, n_steps, n_episodes, auto_reset, t, n_threads, actions,
                           s_y_start, knots, obs_out, reward_out, done_out, term_out, ep_len_out,
                           final_state_out, 0, 0 };
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, rollout_worker, &jobs[t]);
    }
    rollout_worker(&jobs[0]);
    long long total = jobs[0].total;
    int err = jobs[0].err;
    for (int t = 1; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].total;
        if (jobs[t].err) err = jobs[t].err;
    }
    free(jobs);
    free(th);
    return err ? (long long)err : total;
}

/* ------------------------------------------------------------------------- */
/* toy_car.py:7-33                                                           */
/* ------------------------------------------------------------------------- */
void oracle_toy_car(double accel, double v_limit, double dtheta, double dt, int n_iter,
                    double *traj, double *out2)
{
    integrator a_i, vx_i, vy_i;
    integrator_init(&a_i, 0.0, -INFINITY, v_limit, dt);   /* :11, default dt 0.1 */
    integrator_init(&vx_i, 0.0, -INFINITY, INFINITY, dt); /* :12 */
    integrator_init(&vy_i, 0.0, -INFINITY, INFINITY, dt); /* :13 */
    double car_angle = 0.0, s_x = 0.0, s_y = 0.0;
    for (int k = 0; k < n_iter; ++k) {
        car_angle += dtheta;                        /* :23 */
        double v = integrate_signal(&a_i, accel);   /* :24 */
        double v_x = v * cos(car_angle);            /* :26 */
        double v_y = v * sin(car_angle);            /* :27 */
        s_x = integrate_signal(&vx_i, v_x);         /* :29 */
        s_y = integrate_signal(&vy_i, v_y);         /* :30 */
        if (traj) { traj[2 * k] = s_x; traj[2 * k + 1] = s_y; }
    }
    out2[0] = s_x;
    out2[1] = s_y;
}

/* ------------------------------------------------------------------------- */
/* toy_parachute.py:8-41                                                     */
/* ------------------------------------------------------------------------- */
int oracle_toy_parachute(double h0, double h1, double area_free, double area_chute, double mass,
                         double c_w, double rho, double g, double dt_integrator, int max_iter,
                         double *traj, double *out_sv)
{
    integrator a_i, v_i;
    integrator_init(&a_i, 0.0, -INFINITY, INFINITY, dt_integrator); /* :18 */
    integrator_init(&v_i, h0, -INFINITY, INFINITY, dt_integrator);  /* :19 */
    double total_a = 0.0, v = 0.0, s = h0;
    int calls = 0;
    for (int k = 0; k < max_iter; ++k) {
        total_a -= g;                           /* :24 */
        v = integrate_signal(&a_i, total_a);    /* :25 */
        s = integrate_signal(&v_i, v);          /* :26 */
        ++calls;
        if (traj) { traj[2 * k] = s; traj[2 * k + 1] = v; }
        if (s < 0.0) break;                     /* :29-30 */
        double F_w;
        if (s < h1) F_w = v * v * 0.5 * rho * c_w * area_chute; /* :33-34 */
        else        F_w = v * v * 0.5 * rho * c_w * area_free;  /* :35-36 */
        total_a = F_w / mass;                   /* :38 */
    }
    out_sv[0] = s;
    out_sv[1] = v;
    return calls;
}
