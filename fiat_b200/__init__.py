"""fiat_b200 -- B200-native basis tabulation behind FIAT's `FiniteElement.tabulate` contract."""
from .extract import describe_element, UnsupportedElement  # noqa: F401


def __getattr__(name):
    # torch / CUDA are only needed once something is tabulated
    if name in ("tabulate", "tabulate_into", "tabulate_host", "locate_subcells", "Tabulator", "get_tabulator"):
        from . import api as _t
        return getattr(_t, name)
    if name in ("tabulate_sharded", "gather_shards", "shard_range"):
        from . import shard as _s
        return getattr(_s, name)
    raise AttributeError(name)
