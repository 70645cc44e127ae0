"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference checkout is read-only at
/root/reference and does not exist on the GPU box):

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so these files - produced by
the unmodified reference modules plus the three import shims of oracle/reference_shim.py - are
what pins both the CPU oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_*gpu*).
All inputs are seeded; fixtures are small (a few hundred KB in total).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def rand_graph(n, p, gen, weighted=False, self_loops=False, force01=False):
    up = (torch.rand(n, n, generator=gen) < p).float().triu(1)
    if force01:
        up[0, 1] = 1.0
    if weighted:
        wts = torch.tensor([1.0, 1.5, 2.0, 3.0])
        up = up * wts[torch.randint(0, 4, (n, n), generator=gen)]
    A = up + up.t()
    if self_loops:
        A = A + torch.diag((torch.rand(n, generator=gen) < 0.5).float())
    return A


def np_(t):
    return t.detach().cpu().numpy()


def gen_operators(ref):
    gen = torch.Generator().manual_seed(11)
    graphs = {
        "weighted7": rand_graph(7, 0.4, gen, weighted=True),
        "binary9": rand_graph(9, 0.35, gen),
        "loops6": rand_graph(6, 0.5, gen, self_loops=True),
        "dense8_w": rand_graph(8, 0.6, gen, weighted=True, force01=True),
        "single_edge3": torch.tensor([[0., 0., 0.], [0., 0., 2.], [0., 2., 0.]]),
        "empty4": torch.zeros(4, 4),
        "path5": torch.diag(torch.ones(4), 1) + torch.diag(torch.ones(4), -1),
    }
    out = {}
    for name, A in graphs.items():
        V = torch.zeros(A.shape[0], 2)
        out[name + "/A"] = np_(A)
        for J in (1, 2, 3):
            W, WL, Pm, Pd = ref.operators.graph_operators([V, A], J, True)
            out["%s/J%d/W" % (name, J)] = np_(W)
            out["%s/J%d/WL" % (name, J)] = np_(WL)
            if J == 1:
                out[name + "/Pm"] = np_(Pm)
                out[name + "/Pd"] = np_(Pd)
            W2 = ref.operators.graph_operators([V, A], J, False)
            assert torch.equal(W, W2)
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **out)


def make_instances(ref, sizes, J, gen, nfeat=5, weighted=True):
    inst = []
    for k, n in enumerate(sizes):
        A = rand_graph(n, 0.45, gen, weighted=weighted, force01=True)
        x = torch.randn(n, nfeat, generator=gen)
        t = torch.randn(13, generator=gen)
        W, WL, Pm, Pd = ref.operators.graph_operators([x, A], J, True)
        inst.append([x, A, t, W, WL, Pm, Pd])
    return inst


def gen_batch(ref):
    gen = torch.Generator().manual_seed(21)
    inst = make_instances(ref, [5, 8, 3, 6], 2, gen)
    names = ["X", "W", "T", "XL", "WL", "Pm", "Pd", "mask", "mask_lg", "N_batch", "E_batch"]
    res = ref.batching.prepare_batch(inst, 4, 2)
    out = {"n_inst": np.int64(len(inst))}
    for i, (x, A, t, *_rest) in enumerate(inst):
        out["inst%d/x" % i] = np_(x)
        out["inst%d/A" % i] = np_(A)
        out["inst%d/t" % i] = np_(t)
    for n, v in zip(names, res):
        out["out/" + n] = np_(v)
    np.savez_compressed(os.path.join(OUT, "prepare_batch.npz"), **out)


def run_model(ref, kind, order, h, L, J, sizes, seed, dim_out=2):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    inst = make_instances(ref, sizes, J, gen)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = ref.batching.prepare_batch(inst, 0, J)
    if kind == "simple":
        model = ref.model_mnb.GNN_simple(0, h, L, 5, dim_out, J)
    else:
        model = ref.model_mnb.GNN_lg(0, h, L, 5, dim_out, J, order)
    model.train()
    X.requires_grad = True
    out = {"h": np.int64(h), "L": np.int64(L), "J": np.int64(J), "order": np.int64(order),
           "dim_out": np.int64(dim_out), "n_inst": np.int64(len(inst))}
    for i, (x, A, t, *_rest) in enumerate(inst):
        out["inst%d/x" % i] = np_(x)
        out["inst%d/A" % i] = np_(A)
    for k, v in model.state_dict().items():
        out["param/" + k] = np_(v)
    if kind == "simple":
        state = [X, W]
        l0 = model.layer0(state, N_batch, mask)
        out["layer0/X"] = np_(l0[0])
        y = model(state, N_batch, mask)
    else:
        state = [X, XL, W, WL, Pm, Pd]
        l0 = model.layer0(state, N_batch, mask, E_batch, mask_lg)
        out["layer0/X"] = np_(l0[0])
        out["layer0/XL"] = np_(l0[1])
        y = model(state, N_batch, mask, E_batch, mask_lg)
    G = torch.randn(y.shape, generator=gen)
    (y * G).sum().backward()
    out["out"] = np_(y)
    out["gout"] = np_(G)
    out["grad/X"] = np_(X.grad)
    for k, v in model.named_parameters():
        out["grad/" + k] = np_(v.grad)
    # running stats after the two train-mode passes above, then an eval pass (test_mnb.py:39)
    for name, mod in model.named_modules():
        if hasattr(mod, "running_mean"):
            out["running/%s.mean" % name] = np_(mod.running_mean)
            out["running/%s.std" % name] = np_(mod.running_std)
    model.eval()
    with torch.no_grad():
        ye = model(state, N_batch, mask) if kind == "simple" else \
            model(state, N_batch, mask, E_batch, mask_lg)
    out["out_eval"] = np_(ye)
    return out


def gen_models(ref):
    cases = {
        "gnn_simple_h3_L4_J2": ("simple", 0, 3, 4, 2, [6, 4, 8], 31),
        "gnn_simple_h2_L3_J1": ("simple", 0, 2, 3, 1, [5, 7], 32),
        "gnn_lg1_h2_L3_J1": ("lg", 1, 2, 3, 1, [5, 7, 4], 33),
        "gnn_lg2_h2_L3_J1": ("lg", 2, 2, 3, 1, [5, 7, 4], 34),
        "gnn_lg3_h2_L3_J1": ("lg", 3, 2, 3, 1, [5, 7, 4], 35),
        "gnn_lg1_h3_L4_J2": ("lg", 1, 3, 4, 2, [6, 5], 36),
    }
    for name, (kind, order, h, L, J, sizes, seed) in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **run_model(ref, kind, order, h, L, J, sizes, seed))


def gen_ops(ref):
    """Stand-alone graph_oper / P_multi / BN on a padded batch."""
    gen = torch.Generator().manual_seed(41)
    inst = make_instances(ref, [6, 4, 7], 2, gen)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = ref.batching.prepare_batch(inst, 0, 2)
    out = {"n_inst": np.int64(len(inst))}
    for i, (x, A, t, *_rest) in enumerate(inst):
        out["inst%d/x" % i] = np_(x)
        out["inst%d/A" % i] = np_(A)
    F = 3
    Xn = torch.randn(3, F, X.shape[2], generator=gen, requires_grad=True)
    Xe = torch.randn(3, F, XL.shape[2], generator=gen, requires_grad=True)
    gop, pmul = ref.layers_mnb.graph_oper(), ref.layers_mnb.P_multi()
    y1, y2 = gop(W, Xn), gop(WL, Xe)
    y3, y4 = pmul(Pm, Xe), pmul(Pd, Xe)
    y5, y6 = pmul(Pm.transpose(2, 1), Xn), pmul(Pd.transpose(2, 1), Xn)
    outs = [y1, y2, y3, y4, y5, y6]
    gs = [torch.randn(y.shape, generator=gen) for y in outs]
    sum((y * g).sum() for y, g in zip(outs, gs)).backward()
    torch.manual_seed(5)
    bn = ref.batch_normalization.BN(F)
    bn.train()
    H = torch.randn(3, F, X.shape[2], generator=gen, requires_grad=True)
    yb = bn(H, N_batch, mask)
    gb = torch.randn(yb.shape, generator=gen)
    (yb * gb).sum().backward()
    out.update({"Xn": np_(Xn), "Xe": np_(Xe), "gXn": np_(Xn.grad), "gXe": np_(Xe.grad),
                "bn/H": np_(H), "bn/weight": np_(bn.weight), "bn/bias": np_(bn.bias),
                "bn/out": np_(yb), "bn/gout": np_(gb), "bn/gH": np_(H.grad),
                "bn/gweight": np_(bn.weight.grad), "bn/gbias": np_(bn.bias.grad),
                "bn/running_mean": np_(bn.running_mean), "bn/running_std": np_(bn.running_std)})
    for i, (y, g) in enumerate(zip(outs, gs)):
        out["y%d" % (i + 1)] = np_(y)
        out["g%d" % (i + 1)] = np_(g)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **out)


def gen_ccn(ref):
    gen = torch.Generator().manual_seed(51)
    out = {}
    F6 = torch.randn(2, 3, 3, 3, 3, 3, generator=gen)
    out["collapse/F"] = np_(F6)
    out["collapse/out"] = np_(ref.contraction.collapse6to3(F6))
    util = ref.utils_ccn.CompnetUtils(False)
    T = torch.randn(4, 4, 4, 3, generator=gen, requires_grad=True)
    adj = torch.rand(4, 4, generator=gen)
    y = util.outer_contract(T, adj)
    g = torch.randn(y.shape, generator=gen)
    (y * g).sum().backward()
    out.update({"contract/T": np_(T), "contract/adj": np_(adj), "contract/out": np_(y),
                "contract/gout": np_(g), "contract/gT": np_(T.grad)})
    for order, cls in ((2, ref.model_ccn.CCN_2D), (1, ref.model_ccn.CCN_1D)):
        torch.manual_seed(60 + order)
        net = cls(3, 2, 2, 2, False)
        for k, v in net.state_dict().items():
            out["ccn%d/param/%s" % (order, k)] = np_(v)
        for gi, n in enumerate((5, 6)):
            A = rand_graph(n, 0.5, gen, weighted=(gi == 1), force01=True) + torch.eye(n)
            X = torch.randn(n, 3, generator=gen, requires_grad=True)
            net.zero_grad()
            yo = net(X, A)
            go = torch.randn(yo.shape, generator=gen)
            (yo * go).sum().backward()
            pre = "ccn%d/g%d/" % (order, gi)
            out.update({pre + "A": np_(A), pre + "X": np_(X), pre + "out": np_(yo),
                        pre + "gout": np_(go), pre + "gX": np_(X.grad)})
            for k, v in net.named_parameters():
                out[pre + "grad/" + k] = np_(v.grad)
    np.savez_compressed(os.path.join(OUT, "ccn.npz"), **out)


def gen_checkpoint(ref):
    """What the reference's drivers write and read back (functions/logs.py:99-123: whole-module
    ``torch.save(model)``; scripts/main_gnn.py:143-145: ``torch.load(args.model_path)``): a reference
    GNN_lg pickled after one train-mode pass (running statistics set), its state_dict, and the eval-mode
    output the reloaded model must reproduce."""
    gen = torch.Generator().manual_seed(71)
    torch.manual_seed(71)
    inst = make_instances(ref, [6, 4, 7], 1, gen)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = ref.batching.prepare_batch(inst, 0, 1)
    model = ref.model_mnb.GNN_lg(0, 2, 3, 5, 2, 1, 1)
    model.train()
    state = [X, XL, W, WL, Pm, Pd]
    with torch.no_grad():
        y_train = model(state, N_batch, mask, E_batch, mask_lg)
    for mod in model.modules():        # the reference's running stats are graph-attached plain attributes
        if hasattr(mod, "running_mean"):
            mod.running_mean = mod.running_mean.detach()
            mod.running_std = mod.running_std.detach()
    model.eval()
    with torch.no_grad():
        y_eval = model(state, N_batch, mask, E_batch, mask_lg)
    saved = {k: sys.modules.get(k) for k in ref.sys_modules}
    sys.modules.update(ref.sys_modules)       # pickle looks the classes up as models.gnns.model_mnb.GNN_lg ...
    try:
        torch.save(model, os.path.join(OUT, "ref_gnn_lg_module.pt"))
    finally:
        for k, v in saved.items():
            if v is None:
                del sys.modules[k]
            else:
                sys.modules[k] = v
    torch.save(model.state_dict(), os.path.join(OUT, "ref_gnn_lg_state.pt"))
    out = {"n_inst": np.int64(len(inst)), "out_train": np_(y_train), "out_eval": np_(y_eval)}
    for i, (x, A, t, *_rest) in enumerate(inst):
        out["inst%d/x" % i] = np_(x)
        out["inst%d/A" % i] = np_(A)
    for name, mod in model.named_modules():
        if hasattr(mod, "running_mean"):
            out["running/%s.mean" % name] = np_(mod.running_mean)
            out["running/%s.std" % name] = np_(mod.running_std)
    np.savez_compressed(os.path.join(OUT, "checkpoint.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = reference_shim.load()
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    if only:        # e.g. `python oracle/make_golden.py checkpoint`: regenerate only the named groups
        for name in sorted(only):
            globals()["gen_" + name](ref)
        return
    gen_operators(ref)
    gen_batch(ref)
    gen_models(ref)
    gen_ops(ref)
    gen_ccn(ref)
    gen_checkpoint(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
