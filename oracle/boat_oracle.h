/*
 * boat_oracle.h -- CPU restatement (plain C, fp64) of the reference's environment
 * step hot path.  TEST INFRASTRUCTURE ONLY: it may be imported, linked or executed
 * only by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
 * legs of bench.py -- never by the product path (sac-agent_b200/), which has no
 * CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against (a) the reference's own recorded fixtures
 * ressources/settings_visualized/experiment_setting_{1..6} (condensed into
 * tests/golden/fixture_exp*.npz by oracle/make_golden.py) and (b) trajectories
 * produced by the unmodified reference classes imported under oracle/ref_shim.py
 * (tests/golden/ref_rollout_exp*.npz), and, while /root/reference is mounted,
 * directly against the live reference (tests/test_oracle_vs_reference.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).
 */
#ifndef BOAT_ORACLE_H
#define BOAT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Config values read by the hot path (configs/original_config.yaml). */
typedef struct {
    int experiment;      /* base_settings.experiment 1..6         wind.py:30        */
    int test_mode;       /* base_settings.test_mode               boat_env.py:72    */
    double dt;           /* base_settings.dt                      boat_env.py:153   */
    double t_max;        /* base_settings.t_max                   boat_env.py:154   */
    double track_width;  /* boat_env.track_width                  boat_env.py:148   */
    double oob_offset;   /* boat_env.boat_out_of_bounds_offset    boat_env.py:200   */
    double goal_line;    /* boat_env.goal_line                    boat_env.py:85    */
    double fuel;         /* boat.fuel                             boat_env.py:180   */
    double boat_m, boat_m_x, boat_m_y, boat_I, boat_Iz;
    double propeller_diameter, wake_friction, c_r_front, c_r_side, thrust_deduction;
    double rho, boat_area_front, boat_area_side, boat_l, boat_b, rudder_area;
    int fixed_points;    /* wind.fixed_points                     wind.py:73        */
    double max_velocity; /* wind.max_velocity                     wind.py:41        */
    double direction;    /* wind.direction (degrees)              wind.py:44        */
} oracle_params;

enum { ORACLE_TERM_NONE = 0, ORACLE_TERM_GOAL = 1, ORACLE_TERM_OOB = 2, ORACLE_TERM_FUEL = 3,
       ORACLE_TERM_TIMEOUT = 4, ORACLE_TERM_RUDDER = 5 };

typedef struct oracle_env oracle_env;

/* wind.py:12-16: length of the per-episode wind tables, int(t_max / dt). */
int oracle_wind_length(const oracle_params *p);

/* wind.py:69-90 generate_random_curve with the np.random.sample() draw supplied by
 * the caller: not-a-knot cubic through n_knots uniform knots on [0, L], sampled at
 * linspace(0, L, num=L), min/max renormalised iff any sample is outside [0, 1].
 * Returns 0, or -1 if n_knots < 4 (the reference's ValueError, wind.py:73-75). */
int oracle_random_curve(const double *knots, int n_knots, int L, double *out);

/* wind.py:26-67 generate_wind.  knots_a / knots_b are the first / second
 * np.random.sample draws of the episode (only the experiments that draw use them).
 * Returns 0, -1 for fixed_points < 4, -2 for an unknown experiment (wind.py:65-67). */
int oracle_generate_wind(const oracle_params *p, const double *knots_a, const double *knots_b,
                         double *wind_velocity, double *wind_angle);

oracle_env *oracle_env_create(const oracle_params *p);
void oracle_env_destroy(oracle_env *e);

/* boat_env.py:120-126 reset -> Boat.__init__ :144-201.  s_y_start is the
 * np.random.randint draw (:147), knots_* the np.random.sample draws.  obs[11]. */
int oracle_env_reset(oracle_env *e, int s_y_start, const double *knots_a, const double *knots_b,
                     double *obs);
/* Same, but with explicit wind tables (fixture replay from wind.csv). */
int oracle_env_reset_tables(oracle_env *e, int s_y_start, const double *wind_velocity,
                            const double *wind_angle, double *obs);

/* boat_env.py:67-115 step.  Returns 0, or -3 when stepped past the wind table
 * (the reference's IndexError). */
int oracle_env_step(oracle_env *e, double action, double *obs, double *reward, int *done,
                    int *term_code);

/* boat_env.py:128-140 return_all_data (8 floats: s_x, s_y, v_x, v_y, s_r, action,
 * reward, rudder) -- 'n' is the constant 20. */
void oracle_env_all_data(const oracle_env *e, double *out8);
const double *oracle_env_wind_velocity(const oracle_env *e);
const double *oracle_env_wind_angle(const oracle_env *e);
double oracle_env_episode_reward(const oracle_env *e);

/* Batched roll-out used by the parity tests and the CPU baseline: n_envs
 * independent envs, n_steps steps each, pthreads over envs (n_threads <= 0: all cores).
 *   actions      [n_steps][n_envs] (float32-representable values)
 *   s_y_start    [n_episodes][n_envs]; knots [n_episodes][n_envs][2][fixed_points]
 *                (episode-major: episode e of env i uses row e)
 *   obs_out      [n_steps][n_envs][11] or NULL; reward_out/done_out/term_out
 *                [n_steps][n_envs] or NULL
 *   auto_reset   0 -> keep stepping after done exactly as the reference object
 *                would; 1 -> on done, reset with the next episode's randomness
 *                (error -11 when n_episodes is exhausted)
 *   ep_len_out   [n_envs] steps until the first done (or -1), may be NULL
 * Returns total env-steps executed, or a negative error. */
long long oracle_rollout(const oracle_params *p, int n_envs, int n_steps, int n_episodes,
                         int auto_reset, const double *actions, const int *s_y_start,
                         const double *knots, double *obs_out, double *reward_out,
                         unsigned char *done_out, unsigned char *term_out, int *ep_len_out,
                         double *final_state_out /* [n_envs][8] or NULL */, int n_threads);

/* environment/toy_car.py:7-33.  Runs n_iter loop iterations, writes the (s_x, s_y)
 * of every iteration to traj[n_iter][2] (may be NULL); returns final in out2. */
void oracle_toy_car(double accel, double v_limit, double dtheta, double dt, int n_iter,
                    double *traj, double *out2);
/* environment/toy_parachute.py:8-41.  Returns the number of integrator calls made
 * (2654 for the script constants); traj[max_iter][2] = (s, v) per call, may be NULL. */
int oracle_toy_parachute(double h0, double h1, double area_free, double area_chute, double mass,
                         double c_w, double rho, double g, double dt_integrator, int max_iter,
                         double *traj, double *out_sv);

#ifdef __cplusplus
}
#endif
#endif
