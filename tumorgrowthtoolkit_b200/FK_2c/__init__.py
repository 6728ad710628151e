from .FK_2c import Solver
