"""b200-track: B200-native detect-and-track hot path behind the reference's plugin surface.

The directory name carries a hyphen (task layout); import it with
``importlib.import_module("object-tracking_b200")`` or through the ``object_tracking_b200``
alias package at the repo root.  Sub-packages ``models_detection`` / ``models_tracking`` /
``utility`` mirror the reference's module names so ``trainer.py``-style callers are drop-in
(see INTEGRATION.md).
"""
__version__ = "0.1.0"


def install_reference_aliases() -> None:
    """Make the reference's module names resolve to the B200 plugin classes, so that code written against
    ktzsh/object-tracking (``trainer.py:12-14`` importlib lookup, ``from models_detection.KerasYOLO import
    KerasYOLO``, ``from utility.utils import decode_netout``) runs unchanged.  See INTEGRATION.md section 2."""
    import importlib
    import sys
    pkg = __name__
    for sub, mods in (("models_detection", ("KerasYOLO", "YOLO")),
                      ("models_tracking", ("BaseTracker", "TinyTracker", "TinyHeatmapTracker", "MultiObjDetTracker")),
                      ("utility", ("utils",))):
        sys.modules[sub] = importlib.import_module(f"{pkg}.{sub}")
        for m in mods:
            sys.modules[f"{sub}.{m}"] = importlib.import_module(f"{pkg}.{sub}.{m}")
