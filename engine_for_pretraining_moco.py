"""Drop-in for the reference module of the same name (run_mae_pretraining_moco.py:31 imports train_one_epoch from it)."""
from dig_b200.engine import train_one_epoch  # noqa: F401
