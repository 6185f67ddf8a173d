"""`from pit import *` shim so the reference's train_*.py scripts run unchanged from this directory."""
from position_induced_transformer_b200.pit import *  # noqa: F401,F403
from position_induced_transformer_b200.pit import __all__  # noqa: F401
