"""Drop-in for the reference module of the same name (run_class_finetuning.py imports train_one_epoch and evaluate from it)."""
from dig_b200.engine_finetune import evaluate, train_one_epoch  # noqa: F401
