"""The subset of the ``faiss`` module audio-tokens touches, backed by libat_b200.

    import at_b200.faiss_compat as faiss          # or: sys.modules["faiss"] = at_b200.faiss_compat
    faiss.get_num_gpus(); faiss.Kmeans(d, k, niter=..., verbose=..., gpu=...); faiss.IndexFlatL2(d)

(reference call sites: processors/cluster_creator.py:26,42-58; processors/spec_tokenizer.py:125-126,77).
"""
from . import get_num_gpus  # noqa: F401
from .index import IndexFlatL2  # noqa: F401
from .kmeans import ClusteringParameters, Kmeans  # noqa: F401
