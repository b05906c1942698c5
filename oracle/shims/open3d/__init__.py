"""TEST INFRASTRUCTURE ONLY (oracle shim).  open3d is reached only by the ICP branches
(models/egomotion.py:21-23, models/alignnet.py:79-81; both disabled, configs/default.yaml:116-117)
and by visualisation helpers; any attribute access raises."""


class _Stub:
    def __getattr__(self, name):
        raise RuntimeError("open3d is not available (oracle shim); ICP/visualisation branches are out of scope")


pipelines = _Stub()
geometry = _Stub()
utility = _Stub()
registration = _Stub()
visualization = _Stub()
