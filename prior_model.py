"""Drop-in module name of the reference (`import prior_model`): re-exports recombiner_b200.prior_model.
Prior checkpoints pickle `prior_model.LinearTransform` / `prior_model.Upsample` objects
(main_prior_training.py:334-335), so these top-level names must stay importable."""
from recombiner_b200.prior_model import *  # noqa: F401,F403
import sys as _sys

from recombiner_b200 import prior_model as _impl

# unpickling looks classes up by (module, qualname): make both module paths resolve to one class
LinearTransform = _impl.LinearTransform
Upsample = _impl.Upsample
LinearTransform.__module__ = "prior_model"
Upsample.__module__ = "prior_model"
