"""GPU experiment (torchrun, -DCAL_TIMELINE build): where the data-parallel step's exchange time goes.  CTA 0 of
k_dp_adam stamps %globaltimer after its dependency wait, after the pushes, after the system fence, after the flag
stores, after the peers' flags arrived and at its end (spare words of the exchange region's header); printed per rank
next to the step's other stamps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import cal_b200
from bench import build_batches, model_args
from cuda import cudart

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
same = len(sys.argv) > 1 and sys.argv[1] == "same"
batches, cfg = build_batches("spmotif", 128, 8, 1024, 666 + (0 if same else rank))
torch.manual_seed(666)
net = cal_b200.CausalGCN(10, 4, model_args()).cuda().train()
tr = cal_b200.Trainer(net, cal_b200.batch_caps(batches), process_group=True, collective="peer")
dev = [tr.upload(b) for b in batches]
names = {0: "prep start", 2: "fsg_forward start", 3: "fsg_forward end", 4: "ro_fwd start", 6: "ro_bwd start", 8: "fsg_backward start",
         9: "fsg_backward end", 10: "grad_reduce start", 11: "fsg_forward entry", 12: "prep end"}
dnames = ["dp_adam: gradients complete", "dp_adam: pushed", "dp_adam: system fence done", "dp_adam: flags sent",
          "dp_adam: peers' flags seen", "dp_adam: chunk updated, images written (thread 0)", "dp_adam: parameters stored (thread 0)"]
for rep in range(3):
    n = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):
        tr.step(dev[i % 8], dev[(i + 1) % 8])
    e0.record()
    for i in range(20, 20 + n):
        tr.step(dev[i % 8], dev[(i + 1) % 8])
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    st = tr.eng.region("STATUS", torch.int32).cpu().tolist()[96:109]
    hdr = np.zeros(64, np.int32)
    cudart.cudaMemcpy(hdr.ctypes.data, int(tr.peer.comm.region[rank]), 256, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    ev = [(st[k], names[k]) for k in names] + [(int(hdr[16 + k]), dnames[k]) for k in range(7)]
    ev.sort()
    t0 = [v for v, nm in ev if nm == "fsg_forward entry"][0]
    for r in range(world):
        dist.barrier()
        if r == rank and rep == 2:
            print("rank %d  same_batches %s  step %.2f us" % (rank, same, e0.elapsed_time(e1) * 1e3 / n), flush=True)
            for v, nm in ev:
                print("   %9.2f us  %s" % ((v - t0) / 1e3, nm), flush=True)
dist.destroy_process_group()
