"""MessagePassing base, PyG 1.x hook names (message / update, `x_i`/`x_j`
argument suffix convention, flow source_to_target: j = edge_index[0] is the
source, i = edge_index[1] the target the message is aggregated at)."""
import inspect

import torch
from torch_scatter import scatter_add


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr, self.node_dim = aggr, node_dim
        self.__msg_args__ = list(inspect.signature(self.message).parameters)
        self.__upd_args__ = list(inspect.signature(self.update).parameters)[1:]

    def propagate(self, edge_index, size=None, **kwargs):
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() >= 1 and n is None and v.size(0) != edge_index.size(1):
                n = v.size(0)
        if "x" in kwargs:
            n = kwargs["x"].size(0)
        j, i = edge_index[0], edge_index[1]
        args = []
        for name in self.__msg_args__:
            if name.endswith("_i") and name[:-2] in kwargs:
                t = kwargs[name[:-2]]
                args.append(None if t is None else t.index_select(0, i))
            elif name.endswith("_j") and name[:-2] in kwargs:
                t = kwargs[name[:-2]]
                args.append(None if t is None else t.index_select(0, j))
            elif name == "edge_index_i":
                args.append(i)
            elif name == "edge_index_j":
                args.append(j)
            elif name == "size_i" or name == "size_j":
                args.append(n)
            elif name == "edge_index":
                args.append(edge_index)
            else:
                args.append(kwargs[name])
        out = self.message(*args)
        out = scatter_add(out, i, dim=0, dim_size=n)
        return self.update(out, *[kwargs[a] for a in self.__upd_args__])

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
