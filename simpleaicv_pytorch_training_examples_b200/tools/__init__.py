"""Counterpart of the reference's ``tools/`` for the classification hot path: optimizer /
scheduler / training-mode builders (tools/utils.py), the train / test loops (tools/scripts.py)
and the ``train_classification_model.py --work-dir`` entry point."""
