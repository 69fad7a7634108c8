"""Configuration access for the drop-in modules.

The reference exposes one import-time global, ``hparam.hparam`` (hparam.py:47-61), parsed from the CWD-relative
``config/config.yaml``.  When the drop-in runs inside the reference's tree that module is importable and is used
unchanged (we never add keys to config.yaml).  Stand-alone (tests, bench, GPU box) we fall back to the values of
the reference's config/config.yaml (:12, :19-21), overridable through ``configure``.
"""
import sys

DEFAULTS = {"nmels": 40, "hidden": 768, "num_layer": 3, "proj": 256}
_override = {}


def configure(**kw):
    """Override model dimensions for stand-alone use: nmels, hidden, num_layer, proj."""
    for k, v in kw.items():
        if k not in DEFAULTS:
            raise KeyError(k)
        _override[k] = int(v)


def model_dims():
    """(nmels, hidden, num_layer, proj) read exactly like speech_embedder_net.py:19,25 reads hp."""
    dims = dict(DEFAULTS)
    hp_mod = sys.modules.get("hparam")
    hp = getattr(hp_mod, "hparam", None) if hp_mod is not None else None
    if hp is not None:
        try:
            dims.update(nmels=int(hp.data.nmels), hidden=int(hp.model.hidden), num_layer=int(hp.model.num_layer),
                        proj=int(hp.model.proj))
        except (KeyError, AttributeError):
            pass
    dims.update(_override)
    return dims["nmels"], dims["hidden"], dims["num_layer"], dims["proj"]
