/*
 * boatenv.h -- C ABI of libboatenv.so: the B200-native (sm_100a) batched
 * implementation of the environment-step hot path of Nilau1998/SAC-Agent.
 *
 * The reference has no FFI layer: its boundary is the duck-typed Python object that
 * main.py, Recorder and BaseAgent touch.  Every entry point
 * below names the reference interface (file:line, relative to the reference root)
 * whose batched generalisation it is.  The reference-side binding (a ctypes stub)
 * is shown in INTEGRATION.md; sac-agent_b200/_lib.py is the binding this repo ships.
 *
 * Conventions
 *   - plain C types only; every tensor argument is a caller-owned DEVICE pointer
 *     (e.g. torch.Tensor.data_ptr()) unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = OK, negative = argument / configuration error (below),
 *     positive = cudaError_t of the failing CUDA call;
 *   - a handle is bound to one device, is not thread-safe, one handle per GPU;
 *   - there is no CPU fallback: every compute entry point launches sm_100a kernels.
 *   - `precision` is 32 (production: fp32 state/obs, float actions) or 64
 *     (validation: fp64 in the reference's operation order, double actions).
 *     Element type T below means float or double accordingly.  In the 32-bit mode everything a termination
 *     threshold is applied to after accumulation is carried exactly in the same state bytes: the rudder angle
 *     (boat_env.py:72-73, :102-108) as a 44-bit fixed-point number (2^-42 rad), s_x / s_y (:85-93) as int32 fixed
 *     point; get/set_field and env_state_host exchange plain numbers.  Actions are clamped to +-64 before the
 *     rudder update (no outcome changes: |action| >= 21 breaks the rudder from any unbroken state).
 */
#ifndef BOATENV_H
#define BOATENV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BOATENV_API __attribute__((visibility("default")))
#else
#define BOATENV_API
#endif

#define BOATENV_OK 0
#define BOATENV_EINVAL (-1)        /* NULL / out-of-range argument                          */
#define BOATENV_EEXPERIMENT (-2)   /* unknown experiment: the ValueError of wind.py:65-67   */
#define BOATENV_EFIXEDPOINTS (-3)  /* fixed_points < 4: the ValueError of wind.py:73-75     */
#define BOATENV_EUNSUPPORTED (-4)  /* valid for the reference, not for this build (e.g.
                                      fixed_points > 16, t_max/dt > 2^20 - 64)             */
#define BOATENV_ENODEVICE (-5)     /* no CUDA device / not an sm_100 device                 */
#define BOATENV_ESTATE (-6)        /* step() before reset()                                 */
#define BOATENV_EALIGN (-7)        /* a tensor pointer is not 16-byte aligned               */

#define BOATENV_OBS_DIM 11         /* boat_env.py:308-323 return_state                      */

/* Termination codes written to term_out (boat_env.py:84-105, cascade order). */
#define BOATENV_TERM_NONE 0
#define BOATENV_TERM_REACHED_GOAL 1
#define BOATENV_TERM_OUT_OF_BOUNDS 2
#define BOATENV_TERM_OUT_OF_FUEL 3
#define BOATENV_TERM_TIMEOUT 4
#define BOATENV_TERM_RUDDER_BROKEN 5

/* step flags */
#define BOATENV_AUTO_RESET 1u      /* on done: count it, start the next episode in-kernel;
                                      obs_out then holds the NEW episode's first observation
                                      and final_obs_out (if given) the terminal one          */

/* The keys of configs/original_config.yaml that the hot path reads (SURVEY.md
 * section 5).  Plain doubles/ints; filled by the host binding from the YAML. */
typedef struct boatenv_params {
    int32_t experiment;        /* base_settings.experiment  wind.py:30, boat_env.py:166 */
    int32_t test_mode;         /* base_settings.test_mode   boat_env.py:72              */
    double dt;                 /* base_settings.dt          boat_env.py:153             */
    double t_max;              /* base_settings.t_max       boat_env.py:154, wind.py:14 */
    double track_width;        /* boat_env.track_width      boat_env.py:148,311         */
    double oob_offset;         /* boat_env.boat_out_of_bounds_offset  boat_env.py:200   */
    double goal_line;          /* boat_env.goal_line        boat_env.py:85,310          */
    double fuel;               /* boat.fuel                 boat_env.py:180             */
    double boat_m, boat_m_x, boat_m_y, boat_I, boat_Iz;           /* boat_env.py:214-281 */
    double propeller_diameter, wake_friction, c_r_front, c_r_side, thrust_deduction;
    double rho, boat_area_front, boat_area_side, boat_l, boat_b, rudder_area;
    int32_t fixed_points;      /* wind.fixed_points         wind.py:73,77               */
    int32_t _pad;
    double max_velocity;       /* wind.max_velocity         wind.py:41                  */
    double direction;          /* wind.direction, degrees   wind.py:44                  */
} boatenv_params;

typedef struct boatenv_handle *boatenv_t;

/* ---- life cycle ------------------------------------------------------------------ */

/* BoatEnv(config, experiment)  boat_env.py:10-65, batched over n_envs instances.
 * Episode randomness (the np.random draws of boat_env.py:147 and wind.py:78) comes
 * from Philox4x32-10 keyed by `seed` with counter (env_id_offset + i, episode), so
 * results do not depend on how envs are sharded over GPUs. */
BOATENV_API int boatenv_create(const boatenv_params *params, int64_t n_envs, uint64_t seed,
                   int64_t env_id_offset, int precision, int device, boatenv_t *out);
BOATENV_API int boatenv_destroy(boatenv_t h);

/* ---- gym API --------------------------------------------------------------------- */

/* BoatEnv.reset()  boat_env.py:120-126 for every env: new Boat (:144-201), new Wind
 * (wind.py:12-18).  obs_out: T[n_envs][11] or NULL.  mask: uint8[n_envs] or NULL (all). */
BOATENV_API int boatenv_reset(boatenv_t h, const uint8_t *mask, void *obs_out, void *stream);

/* BoatEnv.step(action)  boat_env.py:67-115 for every env.
 *   actions       T[n_envs]            (action[0] of each env; not clipped, like :73)
 *   obs_out       T[n_envs][11]        normalised state (:308-323)
 *   reward_out    T[n_envs]
 *   done_out      uint8[n_envs], or NULL when term_out is given (done == (term != 0): one byte per
 *                 env-step of HBM writes less)
 *   term_out      uint8[n_envs] or NULL  BOATENV_TERM_* of this step
 *   final_obs_out T[n_envs][11] or NULL  written only where done (terminal observation)
 */
BOATENV_API int boatenv_step(boatenv_t h, const void *actions, void *obs_out, void *reward_out,
                 uint8_t *done_out, uint8_t *term_out, void *final_obs_out, uint32_t flags,
                 void *stream);

/* K fused sub-steps of boat_env.py:67-115 with the env state held in registers:
 *   actions  T[K][n_envs] (action_stride = n_envs) or T[n_envs] repeated K times
 *            (action_stride = 0).
 * An env stops at its first done inside the window.  obs_out is the observation after
 * its last executed sub-step (or the new episode's first one under AUTO_RESET),
 * reward_out the sum over executed sub-steps, steps_out (int32[n_envs] or NULL) their
 * number. */
BOATENV_API int boatenv_step_k(boatenv_t h, const void *actions, int64_t action_stride, int32_t k,
                   void *obs_out, void *reward_out, uint8_t *done_out, uint8_t *term_out,
                   int32_t *steps_out, uint32_t flags, void *stream);

/* The same step through HOST buffers (pinned or pageable): copies actions H2D, steps,
 * copies obs/reward/done D2H, chunked over internal streams so that the copies overlap
 * the kernel.  Starts after the work queued on the legacy default stream and blocks until
 * the results are in host memory.  This is the end-to-end
 * call a CPU-side agent loop (main.py:80-81) makes.  Up to 2048 envs (the reference's single-env
 * loop included) it runs zero-copy instead: the kernel reads and writes mapped pinned memory
 * directly, one launch and one synchronize per call. */
BOATENV_API int boatenv_step_host(boatenv_t h, const void *actions_host, void *obs_host, void *reward_host,
                      uint8_t *done_host, uint32_t flags);
/* The same, also returning the termination codes (BOATENV_TERM_*) of this step in term_host: what the
 * single-env drop-in needs to maintain info['termination'] and its counters (boat_env.py:87-105). */
BOATENV_API int boatenv_step_host_term(boatenv_t h, const void *actions_host, void *obs_host, void *reward_host,
                           uint8_t *done_host, uint8_t *term_host, uint32_t flags);
/* boatenv_step_host / _term order themselves after the work queued on the legacy default stream.  This variant
 * names the caller's stream instead (the one its reset / step / learner kernels for this handle were queued on):
 * the internal copy / compute streams wait for an EVENT recorded there -- never a device-wide synchronise, so a
 * learner running on another stream of the same process keeps running.  term_host may be NULL.
 * main.py:80-81 with a CPU-side agent. */
BOATENV_API int boatenv_step_host_stream(boatenv_t h, const void *actions_host, void *obs_host, void *reward_host,
                             uint8_t *done_host, uint8_t *term_host, uint32_t flags, void *stream);
/* K fused sub-steps (boatenv_step_k) through HOST buffers: actions_host T[K][n_envs] in, ONE observation /
 * summed reward / done / termination code / executed-sub-step count per env out -- a host-side agent that
 * repeats or plans K actions moves K times fewer bytes per env-step over PCIe.  term_host and steps_host
 * (int32[n_envs]) may be NULL.  boat_env.py:67-115 K times per call. */
BOATENV_API int boatenv_step_k_host(boatenv_t h, const void *actions_host, int32_t k, void *obs_host, void *reward_host,
                        uint8_t *done_host, uint8_t *term_host, int32_t *steps_host, uint32_t flags, void *stream);

/* ---- state access (env.boat.* of main.py:94, recorder.py:36,46) -------------------- */

enum boatenv_field {
    BOATENV_F_V_X = 0, BOATENV_F_V_Y = 1, BOATENV_F_V_R = 2, BOATENV_F_RUDDER = 3,
    BOATENV_F_S_X = 4, BOATENV_F_S_Y = 5, BOATENV_F_S_R = 6, BOATENV_F_EPISODE_REWARD = 7,
    BOATENV_F_STEP_INDEX = 8,   /* uint32: Boat.index; t = index*dt, fuel = fuel0 - index */
    BOATENV_F_EPISODE = 9       /* uint32: episodes started by this env, minus one        */
};
/* out / in: T[n_envs] for fields 0..7, uint32[n_envs] for fields 8..9 (device). */
BOATENV_API int boatenv_get_field(boatenv_t h, int field, void *out, void *stream);
BOATENV_API int boatenv_set_field(boatenv_t h, int field, const void *in, void *stream);

/* All ten fields of ONE env (boatenv_field order, as doubles) into HOST memory with one launch:
 * what `env.boat.s_x ... rudder_angle` of main.py:94 and `return_all_data` (boat_env.py:128-140,
 * called by the Recorder every step, recorder.py:24-36) read.  Blocking. */
BOATENV_API int boatenv_env_state_host(boatenv_t h, int64_t env_index, double *out_host /* [10] */);

/* env.boat.wind.wind_velocity / .wind_angle (wind.py:16-17, recorder.py:46-47): the
 * current episode's tables of env `env_index`, double[L] each (device). */
BOATENV_API int boatenv_wind_table(boatenv_t h, int64_t env_index, double *wind_velocity_out,
                       double *wind_angle_out, void *stream);
BOATENV_API int boatenv_wind_length(boatenv_t h); /* int(t_max/dt)  wind.py:14-15 */

/* Validation hook: replace the Philox draws by caller-supplied ones for every episode
 * started after this call (fixture replay, differential tests against the reference).
 *   s_y_start int32[n_envs] or NULL; knots double[n_envs][2][fixed_points] or NULL
 * (device; copied).  Passing both NULL restores Philox. */
BOATENV_API int boatenv_set_episode_draws(boatenv_t h, const int32_t *s_y_start, const double *knots,
                              void *stream);

/* HOST function (no GPU work): the draws Philox makes for (seed, global env id,
 * episode): s_y_start as np.random.randint(-0.8*W, 0.8*W) (boat_env.py:147-150) and
 * knots_out[2][fixed_points] as the two np.random.sample(fixed_points) draws
 * (wind.py:78).  Lets a CPU reference be fed the identical episode randomness. */
BOATENV_API int boatenv_episode_draws_host(const boatenv_params *params, uint64_t seed, int64_t global_env_id,
                               uint32_t episode, int32_t *s_y_start_out, double *knots_out);

/* The same for n_ids global env ids and episodes [episode_begin, episode_begin + n_episodes) in one call
 * (the sampled-subset parity tests at benchmark size draw 10^5 episodes):
 *   s_y_start_out int32[n_episodes][n_ids] or NULL; knots_out double[n_episodes][n_ids][2][fixed_points] or NULL. */
BOATENV_API int boatenv_episode_draws_batch_host(const boatenv_params *params, uint64_t seed,
                               const int64_t *global_env_ids, int64_t n_ids, uint32_t episode_begin,
                               int32_t n_episodes, int32_t *s_y_start_out, double *knots_out);

/* ---- checkpoint / resume ----------------------------------------------------------- */

/* The reference checkpoints network weights only (networks/base_network.py:13-17); env state is
 * lost on restart.  Here the complete env state (all per-env scalars, step / episode indices,
 * wind coefficients, the cumulative statistics) is one opaque device blob: together with the
 * counter-based Philox streams it makes resume exact.  boatenv_state_bytes gives its size;
 * export / import copy it to / from a caller-owned device buffer.  A blob is only valid for a
 * handle created with the same params, n_envs and precision. */
BOATENV_API int64_t boatenv_state_bytes(boatenv_t h);
BOATENV_API int boatenv_export_state(boatenv_t h, void *blob_out, void *stream);
BOATENV_API int boatenv_import_state(boatenv_t h, const void *blob_in, void *stream);

/* ---- statistics (info dict of boat_env.py:24-32, cumulative) ----------------------- */

/* out_host[8] = { reached_goal, out_of_bounds, out_of_fuel, timeout, rudder_broken,
 *                 episodes_finished, sum(episode_reward), sum(episode_reward^2) } */
BOATENV_API int boatenv_get_counters(boatenv_t h, double *out_host, void *stream);
/* The same 8 doubles reduced into a device buffer (for an NCCL all-reduce). */
BOATENV_API int boatenv_reduce_counters(boatenv_t h, double *out_device, void *stream);

/* Fill T[n_envs] with the uniform(-1,1) policy "A1" of SURVEY.md 8(d): Philox keyed by
 * (seed, global env id, step_counter).  Benchmark / test input generator. */
BOATENV_API int boatenv_fill_uniform_actions(boatenv_t h, uint64_t step_counter, double scale, void *actions_out,
                                 void *stream);

/* ---- replay buffer (agent/buffer.py:3-35) ------------------------------------------ */

typedef struct boatreplay_handle *boatreplay_t;

/* ReplayBuffer(max_size, input_shape, n_actions)  buffer.py:4-11 (device resident). */
BOATENV_API int boatreplay_create(int64_t max_size, int32_t obs_dim, int32_t n_actions, int precision,
                      int device, boatreplay_t *out);
BOATENV_API int boatreplay_destroy(boatreplay_t r);
/* store_transition (buffer.py:13-22) for a batch of n rows, row i going to slot
 * (mem_cntr + i) % mem_size.  s, s2: T[n][obs_dim]; a: T[n][n_actions]; r: T[n];
 * done: uint8[n]. */
BOATENV_API int boatreplay_store(boatreplay_t r, int64_t n, const void *s, const void *a, const void *rew,
                     const void *s2, const uint8_t *done, void *stream);
/* sample_buffer (buffer.py:24-35): `batch` uniform draws with replacement from
 * [0, min(mem_cntr, mem_size)), Philox keyed by (seed, counter); gathers the five
 * arrays.  idx_out int64[batch] or NULL. */
BOATENV_API int boatreplay_sample(boatreplay_t r, int64_t batch, uint64_t seed, uint64_t counter, void *s_out,
                      void *a_out, void *r_out, void *s2_out, uint8_t *done_out, int64_t *idx_out,
                      void *stream);
/* The gather of buffer.py:29-33 with caller-supplied indices int64[batch] (device). */
BOATENV_API int boatreplay_gather(boatreplay_t r, int64_t batch, const int64_t *idx, void *s_out, void *a_out,
                      void *r_out, void *s2_out, uint8_t *done_out, void *stream);
/* Checkpoint restore: sets the store counter (buffer.py:6 mem_cntr).  The next store goes to slot mem_cntr % mem_size,
 * samples draw from [0, min(mem_cntr, mem_size)). */
BOATENV_API int boatreplay_set_mem_cntr(boatreplay_t r, int64_t mem_cntr);
BOATENV_API int64_t boatreplay_mem_cntr(boatreplay_t r); /* buffer.py:6  */
BOATENV_API int64_t boatreplay_mem_size(boatreplay_t r); /* buffer.py:5  */

/* env.step + agent.remember (main.py:81-88) in ONE kernel: steps every env and writes
 * the transition (obs_inout as s, action, reward, new obs as s', done) straight into
 * the ring.  obs_inout T[n_envs][11] holds the current observations on entry and the
 * next ones on return.  done_flag_mode: 0 = store done (any termination), 1 = store
 * (term == reached_goal) like main.py:83-88. */
BOATENV_API int boatenv_step_store(boatenv_t h, boatreplay_t r, const void *actions, void *obs_inout,
                       void *reward_out, uint8_t *done_out, uint8_t *term_out, int done_flag_mode,
                       uint32_t flags, void *stream);

/* ---- the actor's tanh-squashed Gaussian head (networks/networks.py:47-70) ------------ */

/* ActorNetwork.sample_normal after the two linear heads, fused: from mean, raw_std, eps (all
 * float32 [rows][n_actions], device; eps = the standard-normal draw) and max_action float32
 * [n_actions] to action [rows][n_actions] and log_prob [rows] (summed over the actions).
 * backward: gradients w.r.t. mean and raw_std of the REPARAMETERISED draw (rsample, :61), given
 * grad_action [rows][n_actions] and / or grad_log_prob [rows] (either may be NULL = zero). */
BOATENV_API int boatagent_gaussian_head_forward(const float *mean, const float *raw_std, const float *eps,
                                    const float *max_action, int64_t rows, int32_t n_actions,
                                    float *action_out, float *log_prob_out, void *stream);
BOATENV_API int boatagent_gaussian_head_backward(const float *mean, const float *raw_std, const float *eps,
                                     const float *max_action, const float *grad_action,
                                     const float *grad_log_prob, int64_t rows, int32_t n_actions,
                                     float *grad_mean_out, float *grad_raw_std_out, void *stream);

/* One optimiser step for ALL of the agent's networks -- torch.optim.Adam with its defaults
 * (networks.py:31,88,121), a learning rate per slot -- and, for slots with a `target`, the Polyak
 * average target = tau * param + (1 - tau) * target of update_network_parameters
 * (continuous_agent.py:66-80), in one launch.  slots_host: HOST array (copied into the kernel
 * arguments), all pointers float32 device tensors of `numel` elements.  state_dev: int64[2] on the
 * device, zero-initialised by the caller: [0] = steps taken (incremented by the launch), [1] internal. */
#define BOATAGENT_ADAM_MAX_SLOTS 64
typedef struct boatagent_adam_slot {
    float *param;
    const float *grad;
    float *exp_avg, *exp_avg_sq;
    float *target;          /* NULL: no target network for this tensor */
    int64_t numel;
    float lr;
    int32_t _pad;
} boatagent_adam_slot;
BOATENV_API int boatagent_adam_polyak_step(const boatagent_adam_slot *slots_host, int32_t n_slots, float beta1,
                               float beta2, float eps, float tau, int64_t *state_dev, void *stream);

/* choose_action (continuous_agent.py:57-61) for n envs in one tcgen05 kernel: ActorNetwork.forward
 * (networks.py:38-45, three 256-wide dense layers, bf16 inputs / fp32 accumulation) and the
 * non-reparameterised tanh-squashed draw (:47-65).  weight_blob: BOATAGENT_POLICY_BLOB_BYTES device
 * bytes, 16-byte aligned: [fc1 256 x 16 (obs_dim zero padded)][fc2 256 x 256] as bf16 in 8x8
 * core-matrix order -- element (row, k) of an R-row matrix at (k / 8) * (R * 16) + row * 16 +
 * (k % 8) * 2 --, then the heads as fp32 [16][256] row-major (mean rows, then std rows, zero
 * padded), then the fp32 biases [256][256][16].  obs float32 [n][obs_dim] (obs_dim <= 16); eps
 * float32 [n][n_actions] standard-normal draws or NULL (then Philox(seed; env, step) + Box-Muller);
 * n_actions 1, 2, 4 or 8. */
#define BOATAGENT_POLICY_BLOB_BYTES (256 * 16 * 2 + 256 * 256 * 2 + 16 * 256 * 4 + 256 * 4 + 256 * 4 + 16 * 4)
BOATENV_API int boatagent_policy_act(const void *weight_blob, const float *obs, const float *eps,
                         const float *max_action, uint64_t seed, uint64_t step, int64_t n, int32_t obs_dim,
                         int32_t n_actions, float *action_out, void *stream);

/* ---- toy integrator envs (environment/toy_car.py, toy_parachute.py) ---------------- */

typedef struct boattoy_handle *boattoy_t;
#define BOATTOY_CAR 0
#define BOATTOY_PARACHUTE 1
/* params_host: double[n_params] shared by all envs, jitter: each env's parameters are
 * scaled by 1 + jitter*u, u = Philox uniform(-1,1) keyed (seed, env, param); env 0 is
 * never jittered.  toy_car params: {accel, v_limit, dtheta, dt}; toy_parachute params:
 * {h0, h1, area_free, area_chute, mass, c_w, rho, g, dt_integrator}. */
BOATENV_API int boattoy_create(int kind, int64_t n_envs, const double *params_host, int32_t n_params,
                   double jitter, uint64_t seed, int precision, int device, boattoy_t *out);
BOATENV_API int boattoy_destroy(boattoy_t t);
BOATENV_API int boattoy_reset(boattoy_t t, void *stream);
/* k loop iterations of toy_car.py:22-32 / toy_parachute.py:23-40 per env.
 * out: T[n_envs][4]: car {s_x, s_y, v, angle}; parachute {s, v, a, integrator calls}.
 * done_out uint8[n_envs] or NULL (parachute: s < 0 reached; the env then stays put). */
BOATENV_API int boattoy_step(boattoy_t t, int32_t k, void *out, uint8_t *done_out, void *stream);
/* HOST function (no GPU work): the per-env parameters that envs [env_begin, env_begin + n_envs) of a toy handle
 * created with (params_host, jitter, seed) use -- out double[n_envs][n_params], env 0 = the script constants
 * (toy_car.py:7-8,11,23; toy_parachute.py:8-15).  Lets a CPU reference run the jittered envs. */
BOATENV_API int boattoy_params_host(int kind, const double *params_host, int32_t n_params, double jitter, uint64_t seed,
                        int64_t env_begin, int64_t n_envs, double *out);

/* ---- misc -------------------------------------------------------------------------- */
BOATENV_API const char *boatenv_version(void);
BOATENV_API const char *boatenv_error_string(int code);
/* Number of kernels this library has launched in this process (all handles). */
BOATENV_API int64_t boatenv_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* BOATENV_H */
