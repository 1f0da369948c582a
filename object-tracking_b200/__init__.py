"""b200-track: B200-native detect-and-track hot path behind the reference's plugin surface.

The directory name carries a hyphen (task layout); import it with
``importlib.import_module("object-tracking_b200")`` or through the ``object_tracking_b200``
alias package at the repo root.  Sub-packages ``models_detection`` / ``models_tracking`` /
``utility`` mirror the reference's module names so ``trainer.py``-style callers are drop-in
(see INTEGRATION.md).
"""
__version__ = "0.1.0"
