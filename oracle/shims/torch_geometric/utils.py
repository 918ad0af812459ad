"""``torch_geometric.utils.scatter`` (PyG 2.5.3) restated: zeros(dim_size).scatter_add_(index, src).
Call sites: ``mace/modules/blocks.py:221-226``, ``mace.py:253-258``, ``nequip/modules/interaction_block.py:96``."""
import torch


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = 0, dim_size=None, reduce: str = "sum") -> torch.Tensor:
    assert reduce in ("sum", "add") and dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    out = src.new_zeros((int(dim_size),) + tuple(src.shape[1:]))
    idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add_(0, idx, src)


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce: str = "sum"):
    """Sort edges by (row, col) and merge duplicates by summation (PyG 2.5.3 coalesce, reduce="sum"): used by the
    bond-edge merge branch of the reference's neighbour list (neighbourhood_list_torch.py:98-156).  num_nodes stays a
    tensor when it is one: the reference wraps this function in torch.jit.trace, which would bake int() of the example
    input into the graph."""
    if num_nodes is None:
        n = edge_index.max() + 1
    else:
        n = num_nodes if torch.is_tensor(num_nodes) else int(num_nodes)
    key = edge_index[0] * n + edge_index[1]
    uniq, inv = torch.unique(key, return_inverse=True)
    out_index = torch.stack([torch.div(uniq, n, rounding_mode="floor"), uniq % n], dim=0)

    def merge(a):
        out = a.new_zeros((uniq.shape[0],) + tuple(a.shape[1:]))
        return out.index_add_(0, inv, a)

    if edge_attr is None:
        return out_index
    if isinstance(edge_attr, (list, tuple)):
        return out_index, [merge(a) for a in edge_attr]
    return out_index, merge(edge_attr)


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce: str = "add"):
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    if edge_attr is None:
        return coalesce(both, None, num_nodes)
    if isinstance(edge_attr, (list, tuple)):
        return coalesce(both, [torch.cat([a, a], dim=0) for a in edge_attr], num_nodes)
    return coalesce(both, torch.cat([edge_attr, edge_attr], dim=0), num_nodes)
