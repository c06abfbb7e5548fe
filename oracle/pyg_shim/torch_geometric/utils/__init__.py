import torch
from torch_scatter import scatter_add, scatter_max


def remove_self_loops(edge_index, edge_attr=None):
    row, col = edge_index
    mask = row != col
    edge_attr = edge_attr if edge_attr is None else edge_attr[mask]
    return edge_index[:, mask], edge_attr


def add_self_loops(edge_index, edge_weight=None, fill_value=1, num_nodes=None):
    loop = torch.arange(0, num_nodes, dtype=torch.long, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((num_nodes,), fill_value)])
    return torch.cat([edge_index, loop], dim=1), edge_weight


def softmax(src, index, num_nodes=None):
    out = src - scatter_max(src, index, dim=0, dim_size=num_nodes)[0][index]
    out = out.exp()
    return out / (scatter_add(out, index, dim=0, dim_size=num_nodes)[index] + 1e-16)


def degree(index, num_nodes=None, dtype=None):
    """torch_geometric.utils.degree (call site feature_expansion.py:102): occurrences of every node id, float."""
    num_nodes = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((num_nodes,), dtype=dtype if dtype is not None else torch.float, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))
