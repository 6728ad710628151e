"""tumorgrowthtoolkit_b200 -- B200-native time stepping behind TumorGrowthToolkit's solver API.

Drop-in import paths (reference: TumorGrowthToolkit/FK/__init__.py:1, FK_2c/__init__.py:1,
FK_DTI/__init__.py:1-2):

    from tumorgrowthtoolkit_b200.FK import Solver
    from tumorgrowthtoolkit_b200.FK_2c import Solver
    from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver
    import tumorgrowthtoolkit_b200.FK_DTI.tools as tools

or `tumorgrowthtoolkit_b200.install_as("TumorGrowthToolkit")` to serve the reference's own
import paths from this package.
"""
import importlib
import sys

__version__ = "0.1.0"


def install_as(name="TumorGrowthToolkit"):
    """Register this package under `name` in sys.modules so that unmodified caller code
    (`from TumorGrowthToolkit.FK_2c import Solver`) runs on the B200 implementation."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    for sub in ("base_solver", "FK", "FK.FK", "FK_2c", "FK_2c.FK_2c", "FK_DTI", "FK_DTI.FK_DTI", "FK_DTI.tools"):
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
    return me
