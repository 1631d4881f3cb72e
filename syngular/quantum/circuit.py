"""Circuit -- gate list runner (reference: quantum/circuit.py:7-50)."""
from syngular.quantum.qbit import Qbit


class Circuit:
    def __init__(self, size, bond=2, initializer="ground", structure=None, chi_max=None, cutoff=0.0):
        self.initializer = initializer
        self.size = size
        self.structure = list(structure) if structure is not None else []
        self.chi_max, self.cutoff = chi_max, cutoff
        self.current_step = 0
        self.current_state = None
        self.states = []
        self.reset()

    def run(self):
        for _ in range(self.current_step, len(self.structure)):
            self.step()

    def reset(self):
        if self.initializer == "ground":
            self.current_state = Qbit(self.size, chi_max=self.chi_max, cutoff=self.cutoff)
            self.states.append(self.current_state)

    def step(self):
        self.current_state @= self.structure[self.current_step]
        self.states.append(self.current_state)
        self.current_step += 1

    def add(self, gate):
        self.structure.append(gate)

    def get(self, index=-1):
        return self.states[index]
