"""GPU experiment (needs a -DCAL_TIMELINE build, CAL_B200_LIB pointing at it): the live schedule of one
graph-replayed training step.  Block 0 of every kernel of the fused small-graph step stamps %globaltimer
after its dependency wait and at its end (status[96 + id]); the stamps of the last replayed step are printed
relative to the first one.  usage: timeline.py [prep-ahead 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cal_b200
from bench import build_batches, model_args
ahead = len(sys.argv) > 1 and sys.argv[1] == "1"
batches, cfg = build_batches("spmotif", 128, 8, 1024, 666)
torch.manual_seed(666)
net = cal_b200.CausalGCN(10, 4, model_args()).cuda().train()
tr = cal_b200.Trainer(net, cal_b200.batch_caps(batches), use_graph=True)
dev = [tr.upload(b) for b in batches]
names = ["prep start", "fsg_prep block 0 images built", "fsg_forward start", "fsg_forward end", "ro_fwd start", "ro_fwd end",
         "ro_bwd start", "ro_bwd end", "fsg_backward start", "fsg_backward end", "grad_reduce start", "fsg_forward entry (CTA 0)", "prep end"]
for rep in range(3):
    n = 40 + rep
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):
        tr.step(dev[i % 8], dev[(i + 1) % 8] if ahead else None)
    e0.record()
    for i in range(20, 20 + n):
        tr.step(dev[i % 8], dev[(i + 1) % 8] if ahead else None)
    e1.record()
    torch.cuda.synchronize()
    st = tr.eng.region("STATUS", torch.int32).cpu().tolist()[96:109]
    order = sorted(range(13), key=lambda k: st[k])
    t0 = min(st)
    print("prep-ahead", ahead, " step %.2f us" % (e0.elapsed_time(e1) * 1e3 / n))
    for k in order:
        if True:
            print("   %8.2f us  %s" % ((st[k] - t0) / 1e3, names[k]))
    if rep == 2:
        d = tr.eng.region("D", torch.int32).cpu()[:2 * 160 * 8].view(2, 160, 8)[:, :128].double()
        f0 = float(st[2])        # forward CTA 0 start stamp
        b0 = float(st[8])
        import numpy as np
        for kern, base, nm in ((0, f0, "k_fsg_forward"), (1, b0, "k_fsg_backward")):
            x = (d[kern] - base) / 1e3
            q = lambda v: "min %.2f  p50 %.2f  p90 %.2f  max %.2f" % (v.min(), v.median(), v.quantile(0.9), v.max())
            print(nm, "per-CTA stamps relative to CTA 0's start (us):")
            print("   after dependency wait :", q(x[:, 0]))
            if kern == 1:
                print("   after 1st all-reduce  :", q(x[:, 4]))
            print("   after last all-reduce :", q(x[:, 1]))
            print("   end of work           :", q(x[:, 2]))
            print("   after re-arm          :", q(x[:, 3]))
            tail = x[:, 2] - x[:, 1]
            print("   tail (last all-reduce -> end):", q(tail))

# ---- the device-resident epoch path (cal_collate + look-ahead): the same stamps over step_epoch() ----
if len(sys.argv) > 2 and sys.argv[2] == "epoch":
    from cal_b200.data import make_dataset
    ds = make_dataset(2048, seed=666, bias=0.9)
    store = cal_b200.GraphStore(ds, torch.device("cuda:0"))
    torch.manual_seed(666)
    net2 = cal_b200.CausalGCN(10, 4, model_args()).cuda().train()
    tr2 = cal_b200.Trainer(net2, store.caps(128), use_graph=True)
    import numpy as np
    rng = np.random.RandomState(1)
    for rep in range(3):
        order = rng.permutation(len(ds))[:16 * 128]
        tr2.begin_epoch(store, order, 128)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(4):
            tr2.step_epoch()
        e0.record()
        for i in range(11):
            tr2.step_epoch()
        e1.record()
        torch.cuda.synchronize()
        st = tr2.eng.region("STATUS", torch.int32).cpu().tolist()[96:109]
        print("epoch path  step %.2f us  (fused path: %s)" % (e0.elapsed_time(e1) * 1e3 / 11, tr2.fused_small_graphs))
        t0 = st[11]
        for k in sorted(range(13), key=lambda k: st[k]):
            if k != 1:
                print("   %8.2f us  %s" % ((st[k] - t0) / 1e3, names[k]))
        tr2.step_epoch()
        tr2.end_epoch()
