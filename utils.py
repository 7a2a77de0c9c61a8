"""Drop-in module name of the reference (`import utils`): re-exports recombiner_b200.utils.
Prior checkpoints pickle `prior_model.LinearTransform` / `prior_model.Upsample` objects
(main_prior_training.py:334-335), so these top-level names must stay importable."""
from recombiner_b200.utils import *  # noqa: F401,F403
