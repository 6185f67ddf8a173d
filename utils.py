"""`from utils import *` shim so the reference's train_*.py scripts run unchanged from this directory."""
from position_induced_transformer_b200.utils import *  # noqa: F401,F403
from position_induced_transformer_b200.utils import __all__  # noqa: F401
