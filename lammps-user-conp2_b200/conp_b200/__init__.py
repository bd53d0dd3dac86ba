"""conp_b200: host-side mirror of USER-CONP2's ``fix conp``/``conq``/``cond``
+ ``pppm/conp`` interface on top of the sm_100a CUDA library
(``libconp_b200.so``, C ABI in include/conp_b200.h).

The product path lives in ``abi`` (ctypes binding) and ``fix_conp`` (hook
order of the reference fix).  ``system``/``mockhost`` stand in for the LAMMPS
host that owns atoms and the PPPM mesh tables.
"""

from .fixargs import FixArgs, FixError, parse_fix_args  # noqa: F401
from .system import System, load_reference_case, make_workload, read_lammps_data  # noqa: F401
from .mockhost import MockLammps  # noqa: F401
