"""oceanparameterizations.jl_b200 — B200-native NDE column-model hot path of OceanParameterizations.jl.

Holds only what the path needs: csrc/ (CUDA kernels + the C ABI, built into libopz.so) and the
host-side mirror of the reference's interface (api.py).  The directory name contains a dot, so it
is imported through `opz_import.load()` at the repo root (registered as module `opz_b200`)."""
from . import _lib  # noqa: F401  (fails loudly if libopz.so is missing)
from ._lib import *  # noqa: F401,F403
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
from .parallel import shard_range, make_allreduce_comm, sharded_loss_grad, init_library_comm, HostComm  # noqa: F401
