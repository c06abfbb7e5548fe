"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's epoch loops,
``train_causal_epoch`` (train_causal.py:162-200) and ``eval_acc_causal`` (train_causal.py:202-223).

They are the callers of the hot path: ``model(data, eval_random=...)`` -> three log-probability
tensors -> KL(uniform || c) + NLL(o) + NLL(co) -> ``loss.backward()`` -> ``optimizer.step()``.
``tests/test_reference_loop.py`` pins this restatement to the reference's own function bodies
(extracted from /root/reference/train_causal.py where that tree exists) and then drives it with the
CUDA modules, which is how "main_syn.py calls it unchanged" is exercised."""
import torch
import torch.nn.functional as F


def _num_graphs(data):
    """utils.py:12-16."""
    return data.num_graphs if data.batch is not None else data.x.size(0)


def train_causal_epoch(model, optimizer, loader, device, args):
    """train_causal.py:162-200 -> (loss, c_loss, o_loss, co_loss, train accuracy of the o head), all
    averaged over len(loader.dataset)."""
    model.train()
    sums = [0.0, 0.0, 0.0, 0.0]
    hits = 0
    for data in loader:
        optimizer.zero_grad()
        data = data.to(device)
        y = data.y.view(-1)
        c_logs, o_logs, co_logs = model(data, eval_random=args.with_random)            # :177
        uniform = torch.ones_like(c_logs, dtype=torch.float).to(device) / model.num_classes
        parts = [None, F.kl_div(c_logs, uniform, reduction="batchmean"),               # :180-182
                 F.nll_loss(o_logs, y), F.nll_loss(co_logs, y)]
        parts[0] = args.c * parts[1] + args.o * parts[2] + args.co * parts[3]          # :183
        hits += o_logs.max(1)[1].eq(y).sum().item()
        parts[0].backward()
        g = _num_graphs(data)
        for i in range(4):
            sums[i] += parts[i].item() * g
        optimizer.step()
    n = len(loader.dataset)
    return sums[0] / n, sums[1] / n, sums[2] / n, sums[3] / n, hits / n


def eval_acc_causal(model, loader, device, args):
    """train_causal.py:202-223 -> (acc_co, acc_c, acc_o)."""
    model.eval()
    hits = [0, 0, 0]                                     # co, c, o
    for data in loader:
        data = data.to(device)
        with torch.no_grad():
            c_logs, o_logs, co_logs = model(data, eval_random=args.eval_random)
        y = data.y.view(-1)
        for i, logs in enumerate((co_logs, c_logs, o_logs)):
            hits[i] += logs.max(1)[1].eq(y).sum().item()
    n = len(loader.dataset)
    return hits[0] / n, hits[1] / n, hits[2] / n
