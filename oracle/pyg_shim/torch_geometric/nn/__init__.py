import torch
import torch.nn.functional as F
from torch.nn import Parameter
from torch_scatter import scatter_add

from .conv import MessagePassing
from .inits import glorot, zeros
from ..utils import add_self_loops, remove_self_loops, softmax


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item()) + 1 if size is None else size
    return scatter_add(x, batch, dim=0, dim_size=size)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item()) + 1 if size is None else size
    s = scatter_add(x, batch, dim=0, dim_size=size)
    c = scatter_add(torch.ones_like(batch, dtype=x.dtype), batch, dim=0, dim_size=size)
    return s / c.clamp(min=1).view(-1, 1)


class GINConv(MessagePassing):
    def __init__(self, nn, eps=0, train_eps=False):
        super().__init__("add")
        self.nn = nn
        self.initial_eps = eps
        if train_eps:                              # PyG 1.x: a Parameter when trained, else a buffer
            self.eps = Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index):
        edge_index, _ = remove_self_loops(edge_index)
        return self.nn((1 + self.eps) * x + self.propagate(edge_index, x=x))


class GATConv(MessagePassing):
    """PyG 1.x GATConv (post-1.3 target-indexed softmax)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0, bias=True):
        super().__init__("add")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.weight = Parameter(torch.Tensor(in_channels, heads * out_channels))
        self.att = Parameter(torch.Tensor(1, heads, 2 * out_channels))
        if bias and concat:
            self.bias = Parameter(torch.Tensor(heads * out_channels))
        elif bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        glorot(self.weight)
        glorot(self.att)
        zeros(self.bias)

    def forward(self, x, edge_index, size=None):
        edge_index, _ = remove_self_loops(edge_index)
        edge_index, _ = add_self_loops(edge_index, num_nodes=x.size(0))
        x = torch.matmul(x, self.weight)
        return self.propagate(edge_index, size=size, x=x)

    def message(self, edge_index_i, x_i, x_j, size_i):
        x_j = x_j.view(-1, self.heads, self.out_channels)
        x_i = x_i.view(-1, self.heads, self.out_channels)
        alpha = (torch.cat([x_i, x_j], dim=-1) * self.att).sum(dim=-1)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = softmax(alpha, edge_index_i, size_i)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        return x_j * alpha.view(-1, self.heads, 1)

    def update(self, aggr_out):
        if self.concat is True:
            aggr_out = aggr_out.view(-1, self.heads * self.out_channels)
        else:
            aggr_out = aggr_out.mean(dim=1)
        if self.bias is not None:
            aggr_out = aggr_out + self.bias
        return aggr_out
