"""Stand-in for ``dora.distrib`` (TEST-ONLY): what ``flashy/distrib.py:21`` and
``flashy/logging.py:18`` import.  Rank and world size come from ``flashy_b200.distrib`` so that
virtual ranks are seen correctly."""
from collections import namedtuple

DistribSpec = namedtuple("DistribSpec", "rank world_size local_rank node_rank num_nodes")


def rank() -> int:
    from flashy_b200 import distrib
    return distrib.rank()


def world_size() -> int:
    from flashy_b200 import distrib
    return distrib.world_size()


def init(backend: str = "nccl") -> None:
    from flashy_b200 import distrib
    distrib.init(backend)


def get_distrib_spec() -> DistribSpec:
    return DistribSpec(rank(), world_size(), rank(), 0, 1)
