"""`EmulatorRunner` placeholder (reference: emulator_runner.py:4-42).

In the reference this is the forked worker process whose `_run` loop performs the FiGAR repeat loop for
its slice of environments.  Here that loop is the `k_round` / work-list sequence inside
libmanette_b200.so (see csrc/pool.cu); the class only exists so that
`Runners(tab_rep, EmulatorRunner, emulators, workers, variables)` keeps its signature."""


class EmulatorRunner(object):
    def __init__(self, tab_rep, i, emulators, variables, queue, barrier):
        raise RuntimeError("manette_b200 runs the FiGAR loop on the GPU; EmulatorRunner objects are never created")
