/* manette_b200 -- C ABI of the B200-native environment pool.
 *
 * Drop-in boundary for the reference's env pool:
 *   runners.py:7-50          Runners(tab_rep, EmulatorRunner, emulators, workers, variables)
 *                            start / stop / get_shared_variables / update_environments / wait_updated
 *   emulator_runner.py:19-42 EmulatorRunner._run        (FiGAR repeat loop)
 *   atari_emulator.py:17-136 AtariEmulator               (get_initial_state / next / get_legal_actions)
 *   exploration_policy.py:70-116 ExplorationPolicy.choose_next_actions   (mn_sample_figar)
 *   paac.py:176,180,226-231 + actor_learner.py:108-114   (mn_nstep)
 *   paac.py:79-83,107-112 PAACLearner.update_memory       (observation history ring, mn_config.history)
 *   paac.py:140-205 rollout bookkeeping of PAACLearner.train (mn_rollout_*)
 *
 * Conventions: every call returns 0 on success, <0 on error (mn_last_error() gives the text);
 * no exceptions cross the boundary; all calls on one handle come from one host thread; device
 * work is enqueued on the caller-supplied stream (a cudaStream_t passed as void*, NULL = default).
 * Pointers named *_dev are device pointers, *_host host pointers.  No torch types anywhere.
 */
#ifndef MANETTE_B200_H
#define MANETTE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mn_pool* mn_handle;

/* one contiguous group of environments running the same cartridge */
typedef struct {
  const char* name;          /* rom file stem, e.g. "breakout" (selects the per-game reward/terminal rules) */
  const uint8_t* rom;        /* cartridge image (host) */
  int rom_size;
  int n_envs;                /* environments in this group */
} mn_game;

typedef struct {
  int device;                /* CUDA device ordinal */
  int n_games;
  const mn_game* games;      /* groups are laid out back to back: env ids 0..N-1 */
  int rgb;                   /* args.rgb                    (atari_emulator.py:42-44) */
  int single_life_episodes;  /* args.single_life_episodes   (atari_emulator.py:34,126-130) */
  int random_start;          /* args.random_start           (atari_emulator.py:33,74-77) */
  int random_seed;           /* args.random_seed; ALE seed of env e = random_seed * (env_id_offset + e + 1) (:20) */
  int env_id_offset;         /* global id of local env 0 when the pool is one shard of a multi-GPU job */
  int nb_choices;            /* K = len(tab_rep), 1..32; 0 with tab_rep NULL = plain PAAC ([0]) until mn_set_tab_rep */
  const int* tab_rep;        /* ExplorationPolicy.get_tab_repetitions() (exploration_policy.py:56-62) */
  int envs_per_warp;         /* tuning: 1..32 lanes of a warp that own an environment (0 = automatic: one block per SM) */
  int draw_all_frames;       /* debug: draw the pixels of all four frames of a next(), not only the two pooled ones */
  int no_reset_memo;         /* debug: emulate every get_initial_state() instead of restoring memoised ones */
  int history;               /* H > 0: keep the learner's H-deep observation history (LSTM nets: n_steps = 5, paac.py:107-112) */
} mn_config;

/* device-resident arrays (the reference's five shared variables, paac.py:97-102, + index forms) */
typedef struct {
  int n_envs, num_actions /* A = max over games */, nb_choices, depth /* 1 gray, 3 rgb */;
  uint8_t* states;           /* (N, 84, 84, 4*depth) NHWC, channel c = d*4 + k, k oldest -> newest */
  float* rewards;            /* (N,)  raw reward summed over the macro step */
  float* terminals;          /* (N,)  0.0 / 1.0 */
  float* actions;            /* (N, A) one-hot, written by the learner */
  float* repetitions;        /* (N, K) one-hot, written by the learner (allocation holds K = 32) */
  int32_t* action_idx;       /* (N,)  index form (used instead of the one-hots when use_indices != 0) */
  int32_t* repetition_idx;   /* (N,) */
  int32_t* next_calls;       /* (N,)  number of next() calls each env executed in the last macro step */
  uint8_t* frames;           /* (N, 2, 210, 160) raw palette-index screens of the last two frames */
  uint8_t* ring;             /* (N, 4, 84, 84, depth) observation ring */
  uint8_t* history;          /* (N, H, 84, 84, 4*depth) ring over H of published states, slot mn_history_head() newest; NULL if off */
  int history_depth;
} mn_buffers;

const char* mn_last_error(void);
int mn_create(const mn_config* cfg, mn_handle* out);
int mn_destroy(mn_handle h);
int mn_get_buffers(mn_handle h, mn_buffers* out);
/* Runners(tab_rep, ...) (runners.py:11): the repetition table may arrive after the emulators were built */
int mn_set_tab_rep(mn_handle h, const int* tab_rep, int nb_choices);
/* AtariEmulator.get_legal_actions() of the game of environment `env`; returns count, fills out[<=18] */
int mn_legal_actions(mn_handle h, int env, int32_t* out);

/* N x AtariEmulator.get_initial_state() (paac.py:98): states filled, rewards/terminals zeroed */
int mn_reset_all(mn_handle h, void* stream);
/* Runners.update_environments(): one FiGAR macro step of every environment, asynchronous */
int mn_step_async(mn_handle h, int use_indices, void* stream);
/* Runners.wait_updated(): blocks until the macro step finished; reports sticky CUDA / emulator errors */
int mn_wait(mn_handle h);
/* The reference's shared `states` array lives in host memory (runners.py:30, paac.py:97).  Register the caller's
 * PAGE-LOCKED copy of it (N,84,84,4*depth) u8 and every later mn_reset_all / mn_step_async / mn_env_* writes the states
 * it publishes into that array as well as into the device buffer -- during the step, as environments finish their
 * repeats, not in one copy after it.  The array is complete when mn_wait returns.  NULL unregisters.  The memory must
 * stay allocated while registered.  Blocks until a step in flight has finished. */
int mn_set_host_states(mn_handle h, uint8_t* states_pinned_host);
/* same macro step with HOST arrays, copies inside: actions (N,A) f32, repetitions (N,K) f32 in;
 * states (N,84,84,4*depth) u8, rewards (N,) f32, terminals (N,) f32 out.  Blocking. */
int mn_step_host(mn_handle h, const float* actions_host, const float* repetitions_host, uint8_t* states_host,
                 float* rewards_host, float* terminals_host, void* stream);

/* single-environment facade (test.py:61-109 style callers): blocking */
int mn_env_reset(mn_handle h, int env, void* stream);                       /* get_initial_state() */
int mn_env_next(mn_handle h, int env, int action_index, float* reward, int* terminal, void* stream); /* next(a) */

/* parity taps (blocking device->host copies) */
int mn_get_ram(mn_handle h, int env, uint8_t* out128_host);
int mn_get_screen(mn_handle h, int env, uint8_t* out33600_host);            /* current raw frame, palette indices */
int mn_get_cpu_state(mn_handle h, int env, int32_t* out10_host);            /* A X Y SP PC PS cycles scanlines bank timer */
int mn_get_lives(mn_handle h, int env, int* lives, int* game_over, int* frame_number);
int mn_total_next_calls(mn_handle h, int64_t* out);                         /* since creation */
int mn_memo_stats(mn_handle h, int64_t* out3);          /* get_initial_state() calls restored from the memo, emulated, stored */
int mn_memo_level1_hits(mn_handle h, int64_t* out);     /* of the emulated ones: restored to the end of the reset unit (only the 4 start frames emulated) */
int mn_total_instructions(mn_handle h, int64_t* out);   /* emulated 6502 instructions since creation */
int mn_redo_count(mn_handle h, int64_t* out);   /* units re-run with every frame drawn (exact fallback), since creation */
int mn_palette(uint8_t* gray128_host, uint8_t* rgb128x3_host);
/* diagnostics: a library built with -DMN_CHECK records the first frame-buffer address violation of the picture side
   {code, value, value}; code 0 = none.  Returns -1 in a normal build (no reference counterpart). */
int mn_check_report(unsigned int* out3);
/* diagnostics (pools created with MN_DIAG=2 in the environment): where the 6502 lanes waited for the picture side */
int mn_diag_counters(mn_handle h, unsigned long long* out4);
/* start no-ops of episode `episode` of global environment `global_env` when random_start is on.  The
 * reference draws them from unseeded random.randint(0, 30) (atari_emulator.py:75); here they are a
 * reproducible function so a CPU oracle can be fed the same schedule.  Returns 0..30. */
int mn_start_noops(uint32_t seed, uint32_t global_env, uint32_t episode);

/* K3 stand-alone: max of two raw index frames -> luminance/RGB -> 84x84 nearest.  frames_dev (n,2,210,160),
 * planes_dev (n,84,84,depth) */
int mn_preprocess(const uint8_t* frames_dev, uint8_t* planes_dev, int n, int rgb, void* stream);
/* K4: ExplorationPolicy.choose_next_actions.  mode 0 multinomial, 1 e-greedy, 2 argmax.
 * pi_dev (N,A) f32, rho_dev (N,K) f32; outputs may be NULL */
int mn_sample_figar(const float* pi_dev, const float* rho_dev, int n, int a, int k, int mode, float epsilon,
                    uint64_t seed, uint32_t step, int32_t* action_idx_dev, int32_t* rep_idx_dev,
                    float* action_onehot_dev, float* rep_onehot_dev, void* stream);
/* K5: reward clip + n-step return / advantage.  rewards/terminals/values (T,N) f32, bootstrap (N,) f32 */
int mn_nstep(const float* rewards_dev, const float* terminals_dev, const float* values_dev, const float* bootstrap_dev,
             double gamma, int clip, int t, int n, float* y_dev, float* adv_dev, void* stream);

/* Observation history of PAACLearner (paac.py:79-83 update_memory, :107-112 initialisation, :200-201 zeroing): kept
 * by mn_reset_all / mn_step_async when mn_config.history = H.  The reference shifts an (N,H,...) array every step; here
 * the new state is written into the next slot of a ring by the kernel that publishes it, and an env whose episode
 * ended has its H entries zeroed (newest included, as the reference does).  Reference order memory[e][j], j oldest ->
 * newest, is ring slot (head + 1 + j) % H.  mn_history_gather materialises that order: out_dev (N,H,84,84,4*depth). */
int mn_history_head(mn_handle h, int* head);
int mn_history_gather(mn_handle h, uint8_t* out_dev, void* stream);

/* K6: the per-step bookkeeping of PAACLearner.train (paac.py:173-205) for all environments in one launch */
typedef struct mn_rollout* mn_rollout_handle;
typedef struct {
  int n_envs, max_local_steps /* T */, num_actions /* A */, nb_choices /* K */;
  float* rewards;            /* (T,N) clipped macro-step rewards            (paac.py:180, actor_learner.py:108-114) */
  float* masks;              /* (T,N) 1 - episode_over                      (paac.py:176) */
  int32_t* actions;          /* (T,N) index of the action taken             (paac.py:163) */
  int32_t* repetitions;      /* (T,N) index of the repetition taken         (paac.py:166) */
  double* episode_reward;    /* (N,)  total_episode_rewards                 (paac.py:179) */
  int32_t* episode_steps;    /* (N,)  emulator_steps                        (paac.py:183) */
  float* actions_sum;        /* (N,A) actions_sum                           (paac.py:154,203) */
  uint64_t* action_rep;      /* (A,K) total_action_rep since mn_rollout_begin (paac.py:142,187-189) */
  double* stats;             /* [episodes finished, sum of their rewards, sum of their lengths, min reward, max reward,
                                global_step]: what the episode-statistics all-reduce carries */
  float* finished_reward;    /* (N,)  total reward of the episodes that ended in the last recorded step, env order (paac.py:192) */
  int32_t* finished_steps;   /* (N,)  their lengths (paac.py:193) */
  int32_t* finished_count;   /* (1,) */
} mn_rollout_buffers;
int mn_rollout_create(int device, int n_envs, int max_local_steps, int num_actions, int nb_choices, const int* tab_rep,
                      mn_rollout_handle* out);
int mn_rollout_destroy(mn_rollout_handle r);
int mn_rollout_get_buffers(mn_rollout_handle r, mn_rollout_buffers* out);
/* start of a rollout (paac.py:142): zero the action x repetition histogram */
int mn_rollout_begin(mn_rollout_handle r, void* stream);
/* local step t: rewards/terminals (N,) f32 as published by the pool, the indices the policy chose (N,) i32 */
int mn_rollout_record(mn_rollout_handle r, int t, const float* rewards_dev, const float* terminals_dev,
                      const int32_t* action_idx_dev, const int32_t* rep_idx_dev, int clip, void* stream);

/* per-kernel device timing with CUDA events on the launching stream, for bench.py's roofline object.
 * kinds: 0 = k_round (emulation), 1 = k_push_frames (K3), 2 = k_emit (stack + publish), 3 = other.
 * mn_profile_end synchronises the device and returns summed milliseconds and launch counts (4 each). */
int mn_profile_begin(mn_handle h);
int mn_profile_end(mn_handle h, double* ms_by_kind4, int64_t* launches_by_kind4);
/* kernels launched since creation (bench.py's gpu_launches) */
int mn_launch_count(mn_handle h, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif
