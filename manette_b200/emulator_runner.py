"""`EmulatorRunner` with the reference's constructor and loop (emulator_runner.py:4-42), in process.

In the reference this is a forked worker whose `_run` loop performs the FiGAR repeat loop for its slice of
environments; `Runners` starts W of them.  Here `Runners.update_environments()` runs that loop for every environment at
once on the GPU (the `k_round` / work-list sequence of csrc/pool.cu), so `Runners` never instantiates this class -- a
forked child could not use the parent's CUDA context anyway.  The class is nevertheless a working restatement of the
worker, one environment at a time through `AtariEmulator.next()`: the same `Action` bookkeeping, early exit and
in-loop `get_initial_state()` as emulator_runner.py:24-41.  It serves callers that drive a slice of emulators
themselves (evaluation scripts, debugging) and the parity test that checks the batched macro step against it
(tests/test_gpu_runners.py::test_emulator_runner_loop_equals_the_batched_macro_step)."""
from .exploration_policy import Action


class EmulatorRunner(object):
    def __init__(self, tab_rep, i, emulators, variables, queue, barrier):
        self.id = i
        self.emulators = emulators
        self.variables = variables          # [states, rewards, terminals, actions, repetitions] slices of this worker
        self.queue = queue
        self.barrier = barrier
        self.tab_rep = tab_rep

    # multiprocessing.Process surface the reference's Runners uses (runners.py:33-42): all in process here
    def start(self):
        self.run()

    def join(self, timeout=None):
        return None

    def run(self):
        self._run()

    def _run(self):
        while True:
            instruction = self.queue.get()
            if instruction is None:
                break
            for i, (emulator, action, rep) in enumerate(zip(self.emulators, self.variables[-2], self.variables[-1])):
                macro_action = Action(self.tab_rep, i, action, rep)
                new_s, reward, episode_over = emulator.next(macro_action.current_action)
                self.variables[0][i] = emulator.get_initial_state() if episode_over else new_s
                self.variables[1][i] = reward
                self.variables[2][i] = episode_over
                while macro_action.is_repeated() and not episode_over:
                    new_s, reward, episode_over = emulator.next(macro_action.repeat())
                    self.variables[0][i] = emulator.get_initial_state() if episode_over else new_s
                    self.variables[1][i] += reward
                    self.variables[2][i] = episode_over
                macro_action.reset()
            self.barrier.put(True)
