"""Two-rank checks on real GPUs (run under torchrun by tests/test_gpu_multi.py; needs >= 2 GPUs):

1. sync_bn: with ``model.sync_bn = True`` the ranks normalise with the statistics of the GLOBAL batch - the outputs of
   the local graphs and the rank-summed flat gradient equal those of ONE process running all graphs
   (models/layers/batch_normalization.py:80-93 normalises over the whole batch).
2. the fused peer-memory all-reduce + Adamax (csrc/p2p.cu) leaves the same parameters as NCCL all-reduce + Adamax
   (up to the order of the floating-point sum) and bit-identical parameters on all ranks."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b, floor=0.0):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), floor, 1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    import hgnn_b200  # noqa: F401
    from hgnn_b200 import synth
    from hgnn_b200.dist import FlatParams, FusedAdamax
    from hgnn_b200.functions.batching import prepare_batch
    from hgnn_b200.models.gnns.model_mnb import GNN_lg

    def run(model, inst, G):
        X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
        out = model([X.cuda(), XL.cuda(), W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        for p in model.parameters():
            p.grad = None
        (out * G).sum().backward()
        return out.detach(), torch.cat([p.grad.reshape(-1) for p in model.parameters()])

    # ---- 1. sync_bn
    per = 3
    data = synth.sbm_dataset(per * world, N=70)
    torch.manual_seed(5)
    model = GNN_lg(0, 2, 5, 5, 2, 1, 1).cuda().train()
    for p in model.parameters():
        dist.broadcast(p.data, 0)
    G = torch.randn(per * world, 2, generator=torch.Generator().manual_seed(1)).cuda()
    ref_model = copy.deepcopy(model)
    model.sync_bn = True
    lo = rank * per
    out, g = run(model, data[lo:lo + per], G[lo:lo + per])
    dist.all_reduce(g)                                   # sum over ranks of the rank gradients
    ref_out, ref_g = run(ref_model, data, G)             # one process, the global batch
    e_out = rel(out, ref_out[lo:lo + per])
    e_g = rel(g, ref_g, 1e-3 * float(ref_g.abs().max()))
    # without sync_bn the local statistics differ: make sure the check can fail
    model.sync_bn = False
    out_local, _ = run(model, data[lo:lo + per], G[lo:lo + per])
    e_local = rel(out_local, ref_out[lo:lo + per])
    print("rank %d sync_bn: out rel err %.2e, summed grad rel err %.2e (local-statistics run differs by %.2e)"
          % (rank, e_out, e_g, e_local), flush=True)
    assert e_out < 1e-4 and e_g < 1e-4 and e_local > 1e-3

    # ---- 2. fused peer all-reduce + Adamax vs NCCL all-reduce + Adamax
    torch.manual_seed(7)
    nets = [torch.nn.Linear(37, 11).cuda() for _ in range(2)]
    for n in nets:
        for p in n.parameters():
            dist.broadcast(p.data, 0)
    nets[1].load_state_dict(nets[0].state_dict())
    fps = [FlatParams(n) for n in nets]
    opts = [FusedAdamax(fps[0], lr=1e-2, peer_allreduce=True), FusedAdamax(fps[1], lr=1e-2, peer_allreduce=False)]
    assert opts[0].peers is not None and opts[1].peers is None
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    for step in range(6):
        gs = [torch.randn(p.shape, generator=gen, device="cuda") for p in nets[0].parameters()]
        for fp, opt, net in zip(fps, opts, nets):
            fp.zero_grad()
            for p, g_ in zip(net.parameters(), gs):
                p.grad = g_.clone()
            fp.all_reduce_grad()
            opt.step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    e_p = rel(fps[0].flat, fps[1].flat)
    gathered = [torch.empty_like(fps[0].flat) for _ in range(world)]
    dist.all_gather(gathered, fps[0].flat)
    same = all(torch.equal(t, gathered[0]) for t in gathered)
    print("rank %d peer all-reduce: params vs NCCL rel err %.2e, ranks bit-identical %s, fault %d"
          % (rank, e_p, same, int(opts[0].fault.item())), flush=True)
    assert e_p < 1e-5 and same and int(opts[0].fault.item()) == 0
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
