"""butd_detr_b200 — B200 (sm_100a) forward hot path of BUTD-DETR behind a C-ABI.

    from butd_detr_b200 import BeaUTyDETR          # drop-in nn.Module (eval forward)
    import butd_detr_b200.pointnet2_ext as _ext     # drop-in for the reference's pointnet2._ext
"""
__all__ = ["BeaUTyDETR", "pointnet2_ext", "synth"]


def __getattr__(name):
    if name == "BeaUTyDETR":
        from .model import BeaUTyDETR
        return BeaUTyDETR
    if name in ("pointnet2_ext", "synth", "engine", "build", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
