"""joint-vae_b200: B200-native (sm_100a) implementation of moxime/joint-vae's train / eval hot path.

The directory name is not an importable identifier; load it with `__graft_entry__.load_package()` (registers it as
`jointvae_b200`) or put this directory on sys.path to get the reference's own top-level names (`cvae`, `module.*`).

    jointvae_b200.cvae.ClassificationVariationalNetwork   <- cvae.py of the reference
    jointvae_b200.module.{losses,priors,optimizers,vae_layers}
    jointvae_b200._native                                   <- ctypes binding of libjvae_sm100.so (include/jvae_b200.h)
"""
from . import _native, engine            # noqa: F401
from . import cvae, distributed           # noqa: F401
from .module import losses, priors, optimizers, vae_layers   # noqa: F401
from . import utils                        # noqa: F401

ClassificationVariationalNetwork = cvae.ClassificationVariationalNetwork
