"""Abstract solver base: same contract as the reference's TumorGrowthToolkit/base_solver.py:1-6
(stores the parameter dict by reference, no copy, no validation)."""


class BaseSolver:
    def __init__(self, params):
        self.params = params

    def solve(self):
        raise NotImplementedError("Solve method must be implemented by the subclass.")
