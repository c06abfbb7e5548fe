"""GPU experiment (needs a -DCAL_PHASE_TIMING build): per-phase cycle counts of CTA 0 of
k_conv_fwd (status[16..]) and k_conv_bwd (status[32..]) at cfg-1 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cal_b200
from bench import build_batches, model_args
batches, cfg = build_batches("spmotif", 128, 4, 1024, 666)
torch.manual_seed(666)
net = cal_b200.CausalGCN(10, 4, model_args()).cuda().train()
tr = cal_b200.Trainer(net, cal_b200.batch_caps(batches), use_graph=False)
dev = [tr.upload(b) for b in batches]
for i in range(6):
    tr.step(dev[i % 4])
torch.cuda.synchronize()
st = tr.eng.region("STATUS", torch.int32).cpu().tolist()
names_f = ["prologue", "csr ptr", "csr seg", "gather", "wait W/sync", "gemm", "epilogue", "reduce+finalize"]
names_b = ["prologue", "csr ptr", "stage+gather", "wait W/sync", "gemm dX", "dX store", "dW outer", "dW store", "reduce+finalize"]
print("k_conv_fwd (last launched = MODE 2 masked, no reduce) cycles:", list(zip(names_f, st[16:24])), "sum", sum(st[16:24]))
print("k_conv_bwd (layer 0) cycles:", list(zip(names_b, st[32:41])), "sum", sum(st[32:41]))
names_rf = ["dep wait", "tmem+perm", "bn1 stats", "fc1 stage+issue", "fc1 mma tail", "epi1+bn2", "y2+fc2", "softmax+loss"]
names_rb = ["dep wait", "d logits", "fc2/bn2 bwd + da1", "dW1 stage+issue", "dW1 mma tail", "dW1 store", "dy1 stage+issue",
            "dy1 mma tail", "bn1 bwd + du"]
print("k_readout_tc_fwd (head 0) cycles:", list(zip(names_rf, st[64:72])), "sum", sum(st[64:72]))
print("k_readout_tc_bwd (head 0) cycles:", list(zip(names_rb, st[80:89])), "sum", sum(st[80:89]))
names_fsg = ["dep wait", "plan+CSR+bn_feat", "feat product", "epilogues", "publish", "raw aggregation", "stats wait+finalize",
             "affine+split", "weight wait", "MMA", "node att+stats", "edge att", "masked epilogue+pool"]
if tr.fused_small_graphs:
    print("k_fsg_forward (CTA 0) cycles:", list(zip(names_fsg, st[48:61])), "sum", sum(st[48:61]))
names_fb = ["dep wait", "pooled grad + att-bwd partial sums", "dz operand+issue", "dW (build, MMA, drain)", "MMA tail + gather dots/totals", "d agg tiles", "masked gather rows", "publish",
            "norm bwd", "all-reduce wait", "att bwd rows", "transpose aggregate", "MMA", "stats epilogue", "bn bwd rows + dW drain", "feat bwd"]
if tr.fused_small_graphs:
    print("k_fsg_backward (CTA 0) cycles:", list(zip(names_fb, st[64:80])), "sum", sum(st[64:80]))
print("k_ro_fwd (head 0) cycles:", list(zip(["dep wait", "operand write", "fc1 product", "bn2 finalise", "fc2 loop", "last-block tail", "rows+thread stats", "bn1 finalise", "TMEM epilogue", "fold+h1 store", "softmax+loss"], st[80:91])))
print("k_ro_bwd (head 0, input-gradient CTA) cycles:", list(zip(["dep wait", "d logits", "fc2/bn2 sums", "d a1 operand", "product", "bn1 bwd + du"], st[112:118])))
print("k_ro_bwd (head 0, weight-gradient CTA) cycles:", list(zip(["dep wait", "d logits", "fc2/bn2 sums", "slices+issue", "product tail", "dW1 drain"], st[120:126])))
names_p = ["wait+zero", "edges+counts", "node pass", "scan", "fill", "sort", "write-out"]
print("k_prep_small structure CTA (slice 0) cycles:", list(zip(names_p, st[96:103])), "sum", sum(st[96:103]))
print("k_prep_small statistics CTA 0 [column sums, grid sum]:", st[112:114], " finishing CTA:", st[116:118])
