"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

numpy / pure-Python restatement of the reference's host-side hot path, each piece
citing the reference lines it follows.  The emulator underneath is the C++ oracle
(oracle/liborc.so through the ``ale_python_interface`` shim).

Pinned here against: the PIL-derived resize tables, the
``tab_rep`` tables of README.md:135 / exploration_policy.py:56-62, closed forms of
the n-step recursion, and -- in this container only -- the reference's own
unmodified modules imported through oracle/ref_harness.py (fixtures committed under
tests/golden/).  The emulator layer itself stays "parity unpinned" (see a2600.hpp).
"""
import multiprocessing as mp
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_HERE, "shims") not in sys.path:
    sys.path.insert(0, os.path.join(_HERE, "shims"))

IMG = 84
NR_IMAGES = 4          # atari_emulator.py:11
ACTION_REPEAT = 4      # atari_emulator.py:12
FRAMES_IN_POOL = 2     # atari_emulator.py:14

# PIL Image.resize((84,84), NEAREST) source indices for a 210x160 input
# (= scipy.misc.imresize(..., interp='nearest'), atari_emulator.py:84).  SURVEY.md B.1.
XMAP = np.array([0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 21, 23, 25, 27, 29, 31, 33, 35, 37, 39, 40, 42, 44, 46, 48,
                 50, 52, 54, 56, 58, 60, 61, 63, 65, 67, 69, 71, 73, 75, 77, 79, 80, 82, 84, 86, 88, 90, 92, 94, 96,
                 98, 99, 101, 103, 105, 107, 109, 111, 113, 115, 117, 119, 120, 122, 124, 126, 128, 130, 132, 134,
                 136, 138, 139, 141, 143, 145, 147, 149, 151, 153, 155, 157, 159], dtype=np.int64)
YMAP = np.array([1, 3, 6, 8, 11, 13, 16, 18, 21, 23, 26, 28, 31, 33, 36, 38, 41, 43, 46, 48, 51, 53, 56, 58, 61, 63,
                 66, 68, 71, 73, 76, 78, 81, 83, 86, 88, 91, 93, 96, 98, 101, 103, 106, 108, 111, 113, 116, 118, 121,
                 123, 126, 128, 131, 133, 136, 138, 141, 143, 146, 148, 151, 153, 156, 158, 161, 163, 166, 168, 171,
                 173, 176, 178, 181, 183, 186, 188, 191, 193, 196, 198, 201, 203, 206, 208], dtype=np.int64)


def palettes():
    """(gray[128], rgb[128,3]) of the oracle's NTSC palette, index = TIA colour byte >> 1."""
    import orc_loader
    g = np.zeros(128, np.uint8)
    c = np.zeros((128, 3), np.uint8)
    orc_loader.lib().orc_palette(g.ctypes.data, c.ctypes.data)
    return g, c


def process_frame_pool(pool):
    """atari_emulator.py:79-88: max over the 2 pooled frames, then the nearest gather.
    pool: (2, 210, 160, D) uint8 -> (84, 84, D) uint8."""
    img = np.amax(pool, axis=0)
    return img[YMAP][:, XMAP]


def preprocess_indices(frame_a, frame_b, rgb):
    """Same result starting from the two raw palette-index screens (210,160) the device keeps:
    luminance / RGB lookup (ALE getScreenGrayscale / getScreenRGB), max, gather."""
    gray, col = palettes()
    if rgb:
        a, b = col[frame_a >> 1], col[frame_b >> 1]
    else:
        a, b = gray[frame_a >> 1][..., None], gray[frame_b >> 1][..., None]
    return process_frame_pool(np.stack([a, b]))


class ObsRing(object):
    """environment.py:58-80 ObservationPool: 4-slot ring, emitted oldest->newest with channel
    index c = d*4 + k (pool laid out (84,84,D,4) then reshaped C-order)."""

    def __init__(self, depth):
        self.depth = depth
        self.pool = np.zeros((IMG, IMG, depth, NR_IMAGES), np.uint8)
        self.head = 0

    def push(self, obs):
        self.pool[:, :, :, self.head] = obs
        self.head = (self.head + 1) % NR_IMAGES

    def stacked(self):
        order = [(self.head + i) % NR_IMAGES for i in range(NR_IMAGES)]
        return np.copy(self.pool[:, :, :, order]).reshape(IMG, IMG, self.depth * NR_IMAGES)


class PortAtariEmulator(object):
    """atari_emulator.py:17-136 restated (same constructor arguments, method names and results)."""

    def __init__(self, actor_id, args, noop_schedule=None):
        from ale_python_interface import ALEInterface
        self.ale = ALEInterface()
        self.ale.setInt(b"random_seed", args.random_seed * (actor_id + 1))       # :20
        self.ale.setFloat(b"repeat_action_probability", 0.0)                     # :22
        self.ale.setInt(b"frame_skip", 1)                                        # :25
        self.ale.setBool(b"color_averaging", False)                              # :26
        self.ale.loadROM((args.rom_path + "/" + args.game + ".bin").encode())    # :27-28
        self.legal_actions = self.ale.getMinimalActionSet()                      # :29
        self.lives = self.ale.lives()
        self.random_start = args.random_start
        self.single_life_episodes = args.single_life_episodes
        self.rgb = bool(args.rgb)
        self.depth = 3 if self.rgb else 1
        self.frames = np.zeros((FRAMES_IN_POOL, 210, 160, self.depth), np.uint8)
        self.frame_head = 0
        self.ring = ObsRing(self.depth)
        # the reference draws the start no-ops from unseeded `random` (:75); the port takes an
        # explicit schedule (iterator of ints) so the device path can be fed the same one
        self.noop_schedule = noop_schedule
        self.global_step = 0

    def get_legal_actions(self):
        return self.legal_actions

    def get_noop(self):
        return [1.0, 0.0]

    def _grab(self):
        if self.rgb:
            img = self.ale.getScreenRGB()
        else:
            img = self.ale.getScreenGrayscale()
        self.frames[self.frame_head] = img
        self.frame_head = (self.frame_head + 1) % FRAMES_IN_POOL

    def _action_repeat(self, a):                                                 # :90-100
        reward = 0
        for _ in range(ACTION_REPEAT - FRAMES_IN_POOL):
            reward += self.ale.act(self.legal_actions[a])
        for _ in range(FRAMES_IN_POOL):
            reward += self.ale.act(self.legal_actions[a])
            self._grab()
        return reward

    def _is_terminal(self):                                                      # :126-133
        over = self.ale.game_over()
        if self.single_life_episodes:
            return over or (self.lives > self.ale.lives())
        return over

    def get_initial_state(self):                                                 # :102-110, :70-77
        self.ale.reset_game()
        self.lives = self.ale.lives()
        if self.random_start:
            wait = next(self.noop_schedule) if self.noop_schedule is not None else 0
            for _ in range(wait):
                self.ale.act(self.legal_actions[0])
        for _ in range(NR_IMAGES):
            self._action_repeat(0)
            self.ring.push(process_frame_pool(self.frames))
        if self._is_terminal():
            raise Exception("This should never happen.")
        return self.ring.stacked()

    def next(self, action):                                                      # :112-124
        reward = self._action_repeat(action)
        self.ring.push(process_frame_pool(self.frames))
        terminal = self._is_terminal()
        self.lives = self.ale.lives()
        self.global_step += 1
        return self.ring.stacked(), reward, terminal


def tab_repetitions(max_repetition, nb_choices):
    """exploration_policy.py:56-62."""
    res = [0] * nb_choices
    res[-1] = max_repetition
    for i in range(1, nb_choices - 1):
        res[i] = int(max_repetition / (nb_choices - 1)) * i
    return res


def figar_macro_step(emulator, action_onehot, rep_onehot, tab_rep):
    """emulator_runner.py:24-41 for one environment: 1 + tab_rep[argmax(rep)] calls of next()
    with early exit + in-step reset on terminal.  Returns (state, reward, terminal, n_next)."""
    a = int(np.argmax(action_onehot))
    left = tab_rep[int(np.argmax(rep_onehot))]
    s, r, over = emulator.next(a)
    state = emulator.get_initial_state() if over else s
    total = r
    n = 1
    while left > 0 and not over:
        left -= 1
        s, r, over = emulator.next(a)
        state = emulator.get_initial_state() if over else s
        total += r
        n += 1
    return state, total, over, n


# --------------------------------------------------------------------------- the worker pool
def _pool_worker(tab_rep, emulators, lo, hi, shapes, shms, cmd_q, done_q):
    from multiprocessing import shared_memory
    opened = [shared_memory.SharedMemory(name=n) for n in shms]
    arrs = [np.ndarray(s, dtype=d, buffer=m.buf) for (s, d), m in zip(shapes, opened)]
    states, rewards, terminals, actions, reps, counts = arrs
    while True:
        if cmd_q.get() is None:
            break
        for k, e in enumerate(range(lo, hi)):
            st, r, over, n = figar_macro_step(emulators[k], actions[e], reps[e], tab_rep)
            states[e] = st
            rewards[e] = r
            terminals[e] = over
            counts[e] = n
        done_q.put(True)
    for m in opened:
        m.close()


class PortRunners(object):
    """runners.py:7-50 + emulator_runner.py restated: W forked workers, shared arrays
    [states u8, rewards f32, terminals f32, actions f32 one-hot, repetitions f32 one-hot].
    (States are kept uint8 -- the reference's c_uint blow-up, runners.py:9, is a quirk, not a result.)
    Adds a per-env `next()` counter array so throughput can be reported."""

    def __init__(self, tab_rep, emulators, workers, variables):
        from multiprocessing import shared_memory
        n = len(emulators)
        assert n % workers == 0
        self.workers = workers
        counts = np.zeros(n, np.int32)
        self._shm, self.variables, shapes = [], [], []
        for v in list(variables) + [counts]:
            m = shared_memory.SharedMemory(create=True, size=max(v.nbytes, 1))
            a = np.ndarray(v.shape, dtype=v.dtype, buffer=m.buf)
            a[...] = v
            self._shm.append(m)
            self.variables.append(a)
            shapes.append((v.shape, v.dtype))
        ctx = mp.get_context("fork")
        self.queues = [ctx.Queue() for _ in range(workers)]
        self.barrier = ctx.Queue()
        per = n // workers
        self.procs = [ctx.Process(target=_pool_worker, daemon=True,
                                  args=(tab_rep, emulators[w * per:(w + 1) * per], w * per, (w + 1) * per, shapes,
                                        [m.name for m in self._shm], self.queues[w], self.barrier))
                      for w in range(workers)]

    def start(self):
        for p in self.procs:
            p.start()

    def stop(self):
        for q in self.queues:
            q.put(None)
        for p in self.procs:
            p.join(timeout=5)
        for m in self._shm:
            m.close()
            try:
                m.unlink()
            except FileNotFoundError:
                pass

    def get_shared_variables(self):
        return self.variables[:5]

    def next_counts(self):
        return self.variables[5]

    def update_environments(self):
        for q in self.queues:
            q.put(True)

    def wait_updated(self):
        for _ in range(self.workers):
            self.barrier.get()


# --------------------------------------------------------------------------- n-step returns
def nstep_returns(rewards_raw, terminals, values, bootstrap, gamma, clip=True):
    """paac.py:176,180,226-231 with actor_learner.py:108-114: float64, exactly the reference's
    operation order.  rewards_raw/terminals/values: (T,N); bootstrap: (N,).  Returns y, adv (T,N) f64."""
    T = rewards_raw.shape[0]
    r = np.asarray(rewards_raw, np.float64)
    if clip:
        r = np.where(r > 1.0, 1.0, np.where(r < -1.0, -1.0, r))
    mask = 1.0 - np.asarray(terminals, np.float32).astype(np.float32)   # paac.py:176 (float32 subtraction)
    mask = mask.astype(np.float64)
    v = np.asarray(values, np.float64)
    R = np.asarray(bootstrap, np.float64).copy()
    y = np.zeros_like(r)
    adv = np.zeros_like(r)
    for t in reversed(range(T)):
        R = r[t] + gamma * R * mask[t]
        y[t] = R
        adv[t] = R - v[t]
    return y, adv


# --------------------------------------------------------------------------- FiGAR sampling
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox-4x32-10 (Salmon et al. 2011) on uint32 numpy arrays; the device sampler uses the same
    generator so draws can be compared index-for-index."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & np.uint64(0xFFFFFFFF) for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def uniforms(n, seed, step):
    """Four float32 uniforms in [0,1) per environment: counter = (env, step, 0, 0), key = seed."""
    x = philox4x32(np.arange(n, dtype=np.uint64), np.full(n, step, np.uint64), np.zeros(n, np.uint64),
                   np.zeros(n, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return [(v >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0) for v in x]


def multinomial_choose(probs, u):
    """exploration_policy.py:108-116 as an inverse-CDF draw: p - float32.epsneg, first index whose
    running float32 sum exceeds u; the remainder falls in the last category like numpy's multinomial."""
    p = np.asarray(probs, np.float32) - np.finfo(np.float32).epsneg
    n, k = p.shape
    idx = np.full(n, k - 1, np.int32)
    cum = np.zeros(n, np.float32)
    done = np.zeros(n, bool)
    for j in range(k):
        cum = (cum + p[:, j]).astype(np.float32)
        hit = (~done) & (u < cum)
        idx[hit] = j
        done |= hit
    return idx


def egreedy_choose(probs, u_test, u_pick, eps):
    """exploration_policy.py:96-106."""
    p = np.asarray(probs, np.float32)
    k = p.shape[1]
    greedy = np.argmax(p, axis=1).astype(np.int32)
    rnd = np.minimum((u_pick * np.float32(k)).astype(np.int32), k - 1)
    return np.where(u_test < np.float32(eps), rnd, greedy).astype(np.int32)


def choose_next_actions(pi, rho, mode, seed, step, eps=0.0):
    """exploration_policy.py:70-87: action and repetition sampled independently.
    mode: 0 multinomial, 1 e-greedy, 2 argmax.  Returns (a_idx, r_idx, a_onehot f32, r_onehot f32)."""
    n = pi.shape[0]
    u = uniforms(n, seed, step)
    if mode == 0:
        a, r = multinomial_choose(pi, u[0]), multinomial_choose(rho, u[1])
    elif mode == 1:
        a, r = egreedy_choose(pi, u[0], u[2], eps), egreedy_choose(rho, u[1], u[3], eps)
    else:
        a, r = np.argmax(pi, axis=1).astype(np.int32), np.argmax(rho, axis=1).astype(np.int32)
    return a, r, np.eye(pi.shape[1], dtype=np.float32)[a], np.eye(rho.shape[1], dtype=np.float32)[r]


# ----------------------------------------------------------------------------- rollout bookkeeping (paac.py:79-83,107-205)
class PortRollout(object):
    """The bookkeeping of `PAACLearner.train` restated with the reference's own loops and array types
    (TEST INFRASTRUCTURE: the checker of mn_rollout_record / the observation history ring).

    paac.py:107-112  memory / whole_memory initialisation (LSTM nets, n_steps = 5)
    paac.py:79-83    update_memory
    paac.py:114-127  accumulators and rollout arrays
    paac.py:142-143  per-rollout histogram
    paac.py:149-205  one local step
    """

    def __init__(self, shared_states, emulator_counts, num_actions, tab_rep, max_local_steps, lstm=False, n_steps=5):
        self.emulator_counts = emulator_counts
        self.num_actions = num_actions
        self.tab_rep = list(tab_rep)
        self.total_repetitions = len(self.tab_rep)
        self.max_local_steps = max_local_steps
        self.lstm_bool = lstm
        self.global_step = 0
        self.total_rewards = []
        self.total_steps = []
        if self.lstm_bool:                                                       # :107-112
            self.n_steps = n_steps
            self.memory = np.zeros(([emulator_counts, self.n_steps] + list(shared_states.shape)[1:]), dtype=np.uint8)
            self.whole_memory = np.zeros(([max_local_steps, emulator_counts, self.n_steps] + list(shared_states.shape)[1:]),
                                         dtype=np.uint8)
            for e in range(emulator_counts):
                self.memory[e, -1, :, :, :] = shared_states[e]
        self.emulator_steps = [0] * emulator_counts                              # :116-117
        self.total_episode_rewards = emulator_counts * [0]
        self.actions_sum = np.zeros((emulator_counts, num_actions))             # :119
        self.rewards = np.zeros((max_local_steps, emulator_counts))             # :122
        self.actions = np.zeros((max_local_steps, emulator_counts, num_actions))
        self.repetitions = np.zeros((max_local_steps, emulator_counts, self.total_repetitions))
        self.episodes_over_masks = np.zeros((max_local_steps, emulator_counts))
        self.begin()

    def begin(self):                                                             # :142-143
        self.total_action_rep = np.zeros((self.num_actions, self.total_repetitions))
        self.nb_actions = 0

    @staticmethod
    def rescale_reward(reward):                                                  # actor_learner.py:108-114
        if reward > 1.0:
            reward = 1.0
        elif reward < -1.0:
            reward = -1.0
        return reward

    def update_memory(self, shared_states, t):                                   # :79-83
        self.whole_memory[t] = self.memory
        self.memory[:, :-1, :, :, :] = self.memory[:, 1:, :, :, :]
        self.memory[:, -1, :, :, :] = shared_states

    def before_step(self, t, new_actions, new_repetitions):                      # :154-166
        self.actions_sum += new_actions
        for e in range(self.emulator_counts):
            self.nb_actions += np.argmax(new_repetitions[e]) + 1
        self.actions[t] = new_actions
        self.repetitions[t] = new_repetitions

    def after_step(self, t, new_actions, new_repetitions, shared_states, shared_rewards, shared_episode_over):   # :173-203
        finished = []
        if self.lstm_bool:
            self.update_memory(shared_states, t)
        self.episodes_over_masks[t] = 1.0 - shared_episode_over.astype(np.float32)
        for e, (actual_reward, episode_over) in enumerate(zip(shared_rewards, shared_episode_over)):
            self.total_episode_rewards[e] += actual_reward
            actual_reward = self.rescale_reward(actual_reward)
            self.rewards[t, e] = actual_reward
            self.emulator_steps[e] += self.tab_rep[np.argmax(new_repetitions[e])] + 1
            self.global_step += 1
            a = np.argmax(new_actions[e])
            r = np.argmax(new_repetitions[e])
            self.total_action_rep[a][r] += 1
            if episode_over:
                self.total_rewards.append(self.total_episode_rewards[e])
                self.total_steps.append(self.emulator_steps[e])
                finished.append((self.total_episode_rewards[e], self.emulator_steps[e]))
                self.total_episode_rewards[e] = 0
                self.emulator_steps[e] = 0
                if self.lstm_bool:
                    self.memory[e] = np.zeros(([self.n_steps] + list(shared_states.shape)[1:]), dtype=np.uint8)
                self.actions_sum[e] = np.zeros(self.num_actions)
        return finished
