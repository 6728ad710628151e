from .FK_DTI import FK_DTI_Solver
from .tools import get_RGB
