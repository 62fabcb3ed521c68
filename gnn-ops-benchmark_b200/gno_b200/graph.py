"""Graph construction on the device: the step that precedes the aggregation path in the
reference's synthetic datasets (graph_benchmark/datasets/fakeDatasets.py:238-259,
`get_edge_index`: randint endpoints → remove_self_loops → to_undirected | coalesce, the
torch_geometric.utils functions imported at fakeDatasets.py:14-15).

Each function is the PyG utility of the same name restated over the library's own radix-sort /
unique kernel (`gno_coalesce`): sorting and de-duplicating the edge list is the expensive part;
concatenation and the self-loop mask are plain tensor plumbing.  CUDA tensors only.
"""
import torch

from . import ops


def _num_nodes(edge_index, num_nodes):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0  # host sync, as PyG's maybe_num_nodes


def remove_self_loops(edge_index, edge_attr=None):
    """torch_geometric.utils.remove_self_loops → (edge_index, edge_attr)."""
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """torch_geometric.utils.coalesce: sort by (row, col), merge duplicate edges (edge_attr
    combined with `reduce`).  Returns edge_index, or (edge_index, edge_attr) when edge_attr is given."""
    n = _num_nodes(edge_index, num_nodes)
    index, value = ops.coalesce(edge_index, edge_attr, n, n, reduce)
    return index if edge_attr is None else (index, value)


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """torch_geometric.utils.to_undirected: add the reverse of every edge, then coalesce."""
    n = _num_nodes(edge_index, num_nodes)
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    attr = None if edge_attr is None else torch.cat([edge_attr, edge_attr], dim=0)
    index, value = ops.coalesce(both, attr, n, n, reduce)
    return index if edge_attr is None else (index, value)


def fake_edge_index(num_src_nodes, num_dst_nodes, avg_degree, is_undirected=False, remove_loops=False,
                    device="cuda", generator=None):
    """fakeDatasets.py:238-259 `get_edge_index` on the device: uniform random endpoints
    (num_src_nodes * avg_degree edges), optional self-loop removal, then to_undirected or coalesce."""
    num_edges = num_src_nodes * avg_degree
    row = torch.randint(num_src_nodes, (num_edges,), dtype=torch.int64, device=device, generator=generator)
    col = torch.randint(num_dst_nodes, (num_edges,), dtype=torch.int64, device=device, generator=generator)
    edge_index = torch.stack([row, col], dim=0)
    if remove_loops:
        edge_index, _ = remove_self_loops(edge_index)
    num_nodes = max(num_src_nodes, num_dst_nodes)
    if is_undirected:
        return to_undirected(edge_index, num_nodes=num_nodes)
    return coalesce(edge_index, num_nodes=num_nodes)


def collate(edge_indices, num_nodes):
    """Mini-batch collation of PyG's DataLoader (OpProfiler.py:195-208): shift graph g's node ids
    by the node count of the graphs before it and concatenate.  Returns (edge_index, batch) where
    batch[i] is the graph of node i (the index global_mean_pool scatters over)."""
    dev = edge_indices[0].device
    counts = torch.as_tensor(list(num_nodes), dtype=torch.int64, device=dev)
    offsets = torch.cumsum(counts, 0) - counts
    edge_index = torch.cat([ei + off for ei, off in zip(edge_indices, offsets.unbind(0))], dim=1)
    batch = torch.repeat_interleave(torch.arange(len(edge_indices), device=dev), counts)
    return edge_index, batch
