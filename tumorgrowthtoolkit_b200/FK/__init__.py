from .FK import Solver
