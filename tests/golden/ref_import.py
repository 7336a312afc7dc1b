"""Import helpers for the reference checkout (build container only; /root/reference is absent on the
GPU box).  Used by make_golden.py and by the `-m "not gpu"` pinning tests when the reference is
present.  A 3-line stub satisfies the import-time `cuvs` dependency (svg/kmeans_utils.py:6)."""
import os
import sys
import types

REF = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "svg"))


def import_kmeans_utils():
    if "cuvs" not in sys.modules:
        cuvs = types.ModuleType("cuvs")
        cluster = types.ModuleType("cuvs.cluster")
        km = types.ModuleType("cuvs.cluster.kmeans")
        km.KMeansParams = object
        km.fit = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("cuvs stub"))
        sys.modules.update({"cuvs": cuvs, "cuvs.cluster": cluster, "cuvs.cluster.kmeans": km})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import svg.kmeans_utils as ku

    return ku


def import_placement(model="hyvideo"):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib

    return importlib.import_module(f"svg.models.{model}.placement")
