"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle and the golden fixtures.

Bars (SURVEY section 4 item 3 / BASELINE north_star): graph arrays bit-exact; relocated coordinates
max|d| / max|ref| <= 1e-5 in fp32; parameter gradients <= 1e-4 relative (lin_key.bias is
analytically zero and is checked against an absolute bound instead).
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gad_testutil as util
from g_adaptivity_b200 import GNN, GRAND_plusConv, MeshGraph, synth
from g_adaptivity_b200 import functional as GF
from oracle import gnn_oracle

pytestmark = pytest.mark.gpu

COORD_TOL = 1e-5
GRAD_TOL = 1e-4
NAMES = util.golden_names()


def cuda_model(dataset, opt_cpu, state_dict=None, **extra):
    opt = copy.deepcopy(opt_cpu)
    opt["device"] = "cuda"
    opt.update(extra)
    m = GNN(dataset, opt).to("cuda")
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m


def oracle_model(dataset, opt_cpu, state_dict=None):
    m = gnn_oracle.GNNRef(dataset, copy.deepcopy(opt_cpu))
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m


def grads_of(model):
    return {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}


def check_grads(got, ref, tol=GRAD_TOL):
    assert sorted(k for k in got if "lin_skip" not in k) == sorted(ref)
    scale = max(g.abs().max().item() for n, g in ref.items() if "lin_key.bias" not in n)
    for n, g in ref.items():
        if "lin_key.bias" in n:
            # analytically zero (softmax shift invariance); the reference holds rounding noise here
            assert got[n].abs().max().item() <= 1e-4 * max(scale, 1e-12), n
            continue
        denom = max(g.abs().max().item(), 1e-3 * scale, 1e-12)
        err = (got[n] - g).abs().max().item() / denom
        assert err <= tol, (n, err)


# --------------------------------------------------------------------------------------
# K0: graph builder, bit-exact
# --------------------------------------------------------------------------------------
def _graph_case(data, opt, dim):
    ei = gnn_oracle.filtered_edge_index(data, opt, dim)
    N = data.x_comp.shape[0]
    ds = synth.SyntheticDataset(dim, opt["mesh_dims"])
    m = cuda_model(ds, opt)
    g = m._graph(data, torch.device("cuda"))
    assert g.E == ei.shape[1]
    assert torch.equal(g.edge_index.cpu(), ei)
    rowptr, col, eid = gnn_oracle.csr_by_destination(ei, N)
    assert torch.equal(g.rowptr.cpu(), rowptr)
    assert torch.equal(g.col.cpu(), col)
    assert torch.equal(g.eid.cpu(), eid)
    t_rowptr, t_dst, t_eid = gnn_oracle.csc_by_source(ei, N)
    assert torch.equal(g.t_rowptr.cpu(), t_rowptr)
    assert torch.equal(g.t_dst.cpu(), t_dst)
    # t_slot links every CSC entry to the CSR slot of the same edge
    slot_of_edge = torch.empty(ei.shape[1], dtype=torch.int64)
    slot_of_edge[eid.long()] = torch.arange(ei.shape[1])
    assert torch.equal(g.t_slot.cpu().long(), slot_of_edge[t_eid.long()])
    deg = torch.diff(rowptr)
    assert g.max_in_deg == int(deg.max())
    return g


@pytest.mark.parametrize("name", NAMES)
def test_graph_builder_matches_reference_edge_list(name):
    fx = util.load_golden(name)
    opt = util.fixture_opt(fx)
    data = util.fixture_batch(fx)
    g = _graph_case(data, opt, len(fx["mesh_dims_list"][0]))
    assert torch.equal(g.edge_index.cpu(), fx["edge_index_filtered"])


@pytest.mark.parametrize("mesh_dims,B,over", [
    ((30, 30), 256, {}), ((200,), 512, {}), ((50, 50), 16, {"self_loops": True}),
    ((200, 200), 1, {}), ((17, 17), 5, {"fix_boundary": False}),
])
def test_graph_builder_bit_exact_at_size(mesh_dims, B, over):
    opt = synth.default_opt(mesh_dims, **over)
    data = synth.make_batch(mesh_dims, B, seed=1)
    g = _graph_case(data, opt, len(mesh_dims))
    if mesh_dims == (200, 200):
        assert g.tile_ptr is None          # too large for one CTA -> streaming kernels
    else:
        assert g.tile_ptr is not None


def test_graph_builder_empty_and_ragged_inputs():
    dev = torch.device("cuda")
    # no edges at all: every row empty
    g = MeshGraph.build(torch.zeros((2, 0), dtype=torch.long), 5, device=dev)
    assert g.E == 0 and torch.equal(g.rowptr.cpu(), torch.zeros(6, dtype=torch.int32))
    # all edges masked away, only the appended loops survive
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    m = torch.ones(3, dtype=torch.bool)
    g = MeshGraph.build(ei, 3, masks=(m, None, None), extra_loops=torch.tensor([2, 0]), device=dev)
    assert g.E == 2 and g.edge_index.cpu().tolist() == [[2, 0], [2, 0]]
    # duplicate edges and a hub row keep stable order
    ei = torch.tensor([[3, 1, 3, 2, 1, 0, 3], [0, 0, 0, 0, 0, 0, 0]])
    g = MeshGraph.build(ei, 4, device=dev)
    assert g.col.cpu().tolist() == [3, 1, 3, 2, 1, 0, 3] and g.eid.cpu().tolist() == list(range(7))
    # out-of-range node id is an error, not a crash
    with pytest.raises(ValueError):
        MeshGraph.build(torch.tensor([[0, 7], [1, 0]]), 3, device=dev)


# --------------------------------------------------------------------------------------
# deformer forward / backward against the golden fixtures (reference's own source)
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("force_stream", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_deformer_matches_reference_fixture(name, force_stream):
    fx = util.load_golden(name)
    opt = util.fixture_opt(fx)
    data = util.fixture_batch(fx)
    model = cuda_model(util.fixture_dataset(fx), opt, fx["state_dict"], gad_force_stream=force_stream)
    model.train()
    out = model(data)
    assert out.shape == fx["x_phys"].shape
    err = util.rel_err(out, fx["x_phys"])
    assert err <= COORD_TOL, err
    # attention of the last layer, reference edge order
    conv = model.conv_layers[-1]
    assert torch.equal(conv.stored_ei.cpu(), fx["edge_index_filtered"])
    alpha = conv.stored_alpha.cpu()
    assert alpha.shape == fx["alpha_last"].shape
    assert (alpha - fx["alpha_last"]).abs().max().item() <= 2e-5
    target = data.x_phys.cuda()
    target = target if target.dim() == 2 else target.unsqueeze(-1)
    loss = F.l1_loss(out, target)
    assert abs(loss.item() - fx["loss"]) <= 1e-5 * max(abs(fx["loss"]), 1e-6)
    loss.backward()
    check_grads(grads_of(model), fx["grads"])
    assert model.conv_layers[0].lin_skip.weight.grad is None


# --------------------------------------------------------------------------------------
# BASELINE configs against the oracle on the same seeded inputs
# --------------------------------------------------------------------------------------
def _compare_with_oracle(mesh_dims, B, over=None, burgers=False, backward=True, seed=0, wscale=1.0, **extra):
    over = over or {}
    opt = synth.burgers_opt(mesh_dims, **over) if burgers else synth.default_opt(mesh_dims, **over)
    dim = len(mesh_dims)
    ds = synth.SyntheticDataset(dim, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=seed, burgers=burgers)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if "lin_key" in n or "lin_query" in n:
                p.mul_(wscale)
    model = cuda_model(ds, opt, ref.state_dict(), **extra)
    model.train()
    out = model(data)
    ref_out = ref(data)
    err = util.rel_err(out, ref_out)
    assert err <= COORD_TOL, err
    if backward:
        gnn_oracle.mesh_loss(ref_out, data.x_phys).backward()
        tgt = data.x_phys.cuda()
        F.l1_loss(out, tgt if tgt.dim() == 2 else tgt.unsqueeze(-1)).backward()
        ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
        # bar: 1e-4 against the oracle evaluated in fp64, widened to twice the fp32 oracle's own
        # deviation from fp64 where that is larger (large batches: sums of 10^5 signed terms)
        g64, floor, scale = util.fp64_grads_and_noise_floor(ds, opt, data, ref, ref_grads)
        util.check_grads_conditioned(grads_of(model), g64, floor, scale, tol=GRAD_TOL)
    return model, out, ref_out, data


def test_config1_15x15_single_mesh():
    _compare_with_oracle((15, 15), 1)


@pytest.mark.parametrize("force_stream", [False, True, "csr"])
def test_config2_30x30_batch256_fwd_bwd(force_stream):
    """mesh-resident ELL kernels / streaming ELL ("wide") kernels / CSR streaming kernels."""
    extra = {"gad_force_stream": bool(force_stream)}
    if force_stream == "csr":
        extra["gad_no_wide"] = True
    model, out, ref_out, data = _compare_with_oracle((30, 30), 256, **extra)
    assert model.last_graph.tile_ptr is not None      # the plan exists; force_stream only bypasses it
    assert (model.last_graph.wide_in is not None) == (force_stream is True)

    # size-independent properties on the full batch
    n = 30
    o = out.detach().cpu().view(256, n, n, 2)
    x = data.x_comp.view(256, n, n, 2)
    for (iy, ix) in ((0, 0), (0, n - 1), (n - 1, 0), (n - 1, n - 1)):      # corners are fixed points
        assert torch.equal(o[:, iy, ix], x[:, iy, ix])
    assert o[:, :, 0, 0].abs().max() <= 1e-6 and (o[:, :, -1, 0] - 1).abs().max() <= 1e-6   # sides stay on sides
    assert o[:, 0, :, 1].abs().max() <= 1e-6 and (o[:, -1, :, 1] - 1).abs().max() <= 1e-6


def test_large_mesh_streams_through_wide_rows():
    """A 64x64 mesh does not fit one CTA: with the cluster kernels switched off, forward + backward
    run on the streaming ELL kernels."""
    model, *_ = _compare_with_oracle((64, 64), 3, gad_no_cluster=True)
    g = model.last_graph
    assert g.tile_ptr is None and g.wide_in is not None and g.wide_deg <= 7 and getattr(g, "clf_in", None) is None


@pytest.mark.parametrize("over", [{}, {"ode_method": "rk4", "num_layers": 5}, {"share_conv": False, "learn_step": True}])
def test_large_mesh_forward_on_a_cluster(over):
    """Module-seam forward of meshes beyond one CTA = ONE launch of the cluster-resident kernel
    (backward, where defined, on the streaming ELL kernels from the states it saved)."""
    # gradients only for the shared-weight case: with one weight set PER LAYER the layer-0 gradients of a
    # 64x64 mesh cancel so strongly that the fp32 oracle itself is 1.5e-5 from the fp64 value
    model, *_ = _compare_with_oracle((64, 64), 3, over=over, backward=not over)
    g = model.last_graph
    assert g.tile_ptr is None and g.clf_in is not None and g.clf_C >= 2
    if not over:
        assert g.cl_in is not None       # the backward ran on the cluster-resident kernel too


def test_cluster_backward_equals_streaming_backward():
    """autograd through the module on 64x64 meshes: cluster-resident backward (one launch) against the
    streaming ELL kernels, including the gradient of a learnable step size and of the inputs."""
    import torch.nn.functional as F
    mesh_dims, B = (64, 64), 2
    over = {"learn_step": True}
    opt = synth.default_opt(mesh_dims, **over)
    ds = synth.SyntheticDataset(2, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=12)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    res = []
    for no_cluster in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_cluster=no_cluster)
        model.train()
        d = data.clone().to("cuda:0")
        d.x_comp.requires_grad_(True)
        out = model(d)
        F.l1_loss(out, d.x_phys).backward()
        g = model.last_graph
        assert (g.cl_in is not None) == (not no_cluster)
        res.append(({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, d.x_comp.grad.clone()))
    for n in res[0][0]:
        a, b = res[0][0][n], res[1][0][n]
        assert (a - b).abs().max().item() <= 2e-5 * max(b.abs().max().item(), 1e-12), n
    assert util.rel_err(res[0][1], res[1][1]) <= 1e-5


def test_config3_burgers_1d_200_repeated_calls():
    """Burgers roll-out pattern (src/utils_eval_Burgers.py:269,297): same graph, new uu each call."""
    mesh_dims, B = (200,), 4096          # BASELINE configs[2] at full size: 819 200 nodes per call
    opt = synth.burgers_opt(mesh_dims)
    ds = synth.SyntheticDataset(1, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=0, burgers=True)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    model = cuda_model(ds, opt, ref.state_dict())
    model.eval()
    ref.eval()
    rng = np.random.default_rng(0)
    with torch.no_grad():
        for call in range(4):
            out = model(data)
            ref_out = ref(data)
            assert util.rel_err(out, ref_out) <= COORD_TOL
            assert out.shape == (B * 200, 1)
            # end points of every 1-D mesh are fixed (self-loop only)
            o = out.cpu().view(B, 200)
            assert torch.equal(o[:, 0], data.x_comp.view(B, 200)[:, 0]) and torch.equal(o[:, -1], data.x_comp.view(B, 200)[:, -1])
            x = data.x_comp.view(B, 200).numpy()
            amp, ph = rng.uniform(0.1, 0.3), rng.uniform(0, 1)
            data.uu_tensor = torch.from_numpy((amp * np.exp(-((x - 0.3 - 0.1 * call - 0.05 * ph) ** 2) / 0.01)).astype(np.float32).reshape(-1))
    assert model._graphs.misses == 1 and model._graphs.hits == 3     # topology cached across calls


@pytest.mark.parametrize("mesh_dims", [(15, 15), (7, 7), (200,), (2, 2)])
def test_device_edge_masks_match_reference_semantics(mesh_dims):
    """gad_edge_masks == the three Python loops of firedrake_mesh_to_PyG (src/data.py:465-494), as
    restated by synth._masks_from_sides (the generator of every fixture), bit for bit."""
    from g_adaptivity_b200 import graph as G
    topo = synth.MeshTopology(mesh_dims)
    n = topo.num_nodes
    side_bits = np.zeros(n, dtype=np.uint8)
    if len(mesh_dims) == 2:
        m = mesh_dims[0]
        ids = np.arange(n).reshape(m, m)
        for k, nodes in enumerate((ids[:, 0], ids[:, -1], ids[0, :], ids[-1, :])):     # markers 1..4
            side_bits[nodes] |= np.uint8(1 << k)
    else:
        side_bits[0] |= 1
        side_bits[n - 1] |= 2
    ei = torch.from_numpy(topo.edge_index).cuda()
    tb, tc, db = G.edge_masks(ei, torch.from_numpy(side_bits))
    assert torch.equal(tb.cpu(), torch.from_numpy(topo.to_boundary_edge_mask))
    assert torch.equal(tc.cpu(), torch.from_numpy(topo.to_corner_nodes_mask))
    assert torch.equal(db.cpu(), torch.from_numpy(topo.diff_boundary_edges_mask))
    with pytest.raises(ValueError):
        bad = ei.clone()
        bad[0, 0] = n + 5
        G.edge_masks(bad, torch.from_numpy(side_bits))


def test_graph_cache_finds_fresh_batch_objects_by_content():
    """The loader of the reference yields a new Batch object per iteration (src/run_GNN.py:97-105):
    the graph cache must recognise the topology by content (device fingerprint), and must NOT confuse
    batches whose topology differs."""
    mesh_dims, B = (12, 12), 6
    opt = synth.default_opt(mesh_dims)
    ds = synth.SyntheticDataset(2, mesh_dims)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    model = cuda_model(ds, opt, ref.state_dict())
    model.eval()
    outs = []
    with torch.no_grad():
        for it in range(3):
            data = synth.make_batch(mesh_dims, B, seed=5)         # fresh tensors, same content
            outs.append(model(data))
    assert model._graphs.misses == 1 and getattr(model._graphs, "misses_identity", 0) == 2
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert util.rel_err(outs[0], ref(synth.make_batch(mesh_dims, B, seed=5))) <= COORD_TOL
    # one edge redirected -> different fingerprint -> rebuilt, and the result follows the new topology
    data = synth.make_batch(mesh_dims, B, seed=5)
    keep = ~(data.to_boundary_edge_mask | data.to_corner_nodes_mask | data.diff_boundary_edges_mask)
    e = int(torch.nonzero(keep)[0])
    data.edge_index = data.edge_index.clone()
    data.edge_index[0, e] = data.edge_index[1, e]                  # becomes a self-loop
    with torch.no_grad():
        out2 = model(data)
        assert model._graphs.misses == 2
        assert util.rel_err(out2, ref(data)) <= COORD_TOL
    assert not torch.equal(out2, outs[0])


def test_shared_topology_option_recognises_batches_without_topology_traffic():
    """`gad_shared_topology` (datasets whose samples all live on one mesh, src/data.py:143): a fresh Batch
    of a known shape is served from the cache by shapes + host samples -- no fingerprint pass, no copy of
    edge_index -- and another batch size is another graph."""
    mesh_dims, B = (12, 12), 6
    opt = synth.default_opt(mesh_dims)
    ds = synth.SyntheticDataset(2, mesh_dims)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    model = cuda_model(ds, opt, ref.state_dict(), gad_shared_topology=True)
    model.eval()
    with torch.no_grad():
        for it in range(3):
            data = synth.make_batch(mesh_dims, B, seed=5 + it)     # fresh tensors, new features, same mesh
            assert util.rel_err(model(data), ref(data)) <= COORD_TOL
        assert model._graphs.misses == 1 and model._graphs.shared_hits == 2
        assert getattr(model._graphs, "misses_identity", 0) == 0
        data = synth.make_batch(mesh_dims, B + 1, seed=9)
        assert util.rel_err(model(data), ref(data)) <= COORD_TOL
        assert model._graphs.misses == 2 and model._graphs.shared_hits == 2


def test_inference_session_replays_burgers_rollout():
    """CUDA-graph replay of the deformer call (GNN.inference_session) == the module call, for a
    sequence of uu fields on a fixed 1-D mesh batch."""
    mesh_dims, B = (200,), 256
    opt = synth.burgers_opt(mesh_dims)
    ds = synth.SyntheticDataset(1, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=3, burgers=True)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    model = cuda_model(ds, opt, ref.state_dict())
    model.eval()
    sess = model.inference_session(data)
    x = data.x_comp.view(B, 200).numpy()
    with torch.no_grad():
        for call in range(3):
            data.uu_tensor = torch.from_numpy((0.2 * np.exp(-((x - 0.3 - 0.1 * call) ** 2) / 0.01)).astype(np.float32).reshape(-1))
            got = sess(uu=data.uu_tensor.cuda()).clone()
            want = model(data)
            assert torch.equal(got, want)
            assert util.rel_err(got, ref(data)) <= COORD_TOL


def test_config4_200x200_rk4_64_steps_forward(monkeypatch):
    over = {"ode_method": "rk4", "num_layers": 64}
    model, out, ref_out, data = _compare_with_oracle((200, 200), 1, over=over, backward=False)
    # a single 16-CTA cluster would leave 90 % of the machine idle: the module dispatches the forward to the
    # chain of 256 dependent streaming launches (graph.stream_fwd_preferred) ...
    assert model.last_graph.tile_ptr is None and model.last_graph.clf_C == 16 and model.last_graph.wide_in is not None
    model.eval()
    with torch.no_grad():
        got = model(data)
    assert util.rel_err(got, ref_out) <= COORD_TOL
    # ... and the cluster kernel (ONE launch for all 64 steps) gives the same mesh
    monkeypatch.setenv("GAD_FWD_POLICY", "cluster")
    with torch.no_grad():
        got2 = model(data)
    assert util.rel_err(got2, ref_out) <= COORD_TOL and util.rel_err(got2, got) <= COORD_TOL
    model2, *_ = _compare_with_oracle((200, 200), 1, over=over, backward=False, gad_no_cluster=True)   # 256 launches
    assert getattr(model2.last_graph, "clf_in", None) is None


@pytest.mark.parametrize("mesh_dims,B,burgers,over", [
    ((64, 64), 1, False, {}), ((64, 64), 1, False, {"ode_method": "rk4", "num_layers": 5}),
    ((200, 200), 1, False, {"ode_method": "rk4", "num_layers": 8}),
    ((90, 90), 1, False, {"share_conv": False, "learn_step": True, "num_layers": 6}),
    ((100, 100), 3, False, {"num_layers": 7}), ((20000,), 2, True, {"ode_method": "rk4", "num_layers": 6}),
    ((70, 70), 2, False, {"self_loops": True, "num_layers": 3, "ode_method": "rk4"}),        # degree 7: the widest rows
    ((70, 70), 1, False, {"fix_boundary": False, "num_layers": 5})])
def test_persistent_streaming_forward_equals_the_launch_chain(mesh_dims, B, burgers, over, monkeypatch):
    """Graphs whose nodes fit the GPU's co-resident threads run all steps of the streaming forward in ONE cooperative
    launch (k_wide_persist: state in registers, window in shared memory, halo rows exchanged through L2 as tagged
    words); same arithmetic as the chain of one launch per F-evaluation, so the result is bit-identical -- forward
    values and, where the backward exists (Euler), the gradients computed from the states it saved."""
    from g_adaptivity_b200 import _lib
    lib = _lib.load()
    outs, launches = [], []
    for persist in ("1", "0"):
        monkeypatch.setenv("GAD_WIDE_PERSIST", persist)
        model, out, ref_out, data = _compare_with_oracle(mesh_dims, B, over=over, burgers=burgers, backward=False,
                                                         gad_no_cluster=True)
        assert model.last_graph.wide_in is not None and model.last_graph.wide_reach >= 1
        model.eval()
        with torch.no_grad():
            model(data)
            torch.cuda.synchronize()
            n0 = lib.gad_launch_count()
            o = model(data)
            torch.cuda.synchronize()
            launches.append(lib.gad_launch_count() - n0)
            for _ in range(3):                     # run to run: the exchange has no data-dependent ordering
                assert torch.equal(model(data), o)
        grads = []
        if over.get("ode_method", "euler") == "euler":
            model.train()
            o_train = model(data)
            assert torch.equal(o_train.detach(), o)
            g = torch.autograd.grad(o_train.square().sum(), [p for p in model.parameters() if p.requires_grad],
                                    allow_unused=True)
            grads = [t.clone() for t in g if t is not None]
        outs.append((o.clone(), grads))
    assert torch.equal(outs[0][0], outs[1][0])
    assert len(outs[0][1]) == len(outs[1][1])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)
    assert launches[0] < launches[1], launches          # one launch instead of one per F-evaluation


def test_rk4_mesh_resident_matches_oracle():
    over = {"ode_method": "rk4", "num_layers": 6}
    model, *_ = _compare_with_oracle((20, 20), 7, over=over, backward=False)
    assert model.last_graph.tile_ptr is not None


@pytest.mark.parametrize("mesh_dims,B,burgers,over", [
    ((20, 20), 7, False, {"num_layers": 3}),
    ((30, 30), 5, False, {"num_layers": 2, "share_conv": False}),          # per-step weights, cfg-2-shaped tiles
    ((12, 12), 9, False, {"num_layers": 4, "self_loops": True, "time_step": 0.3}),
    ((40,), 33, True, {"num_layers": 3}),                                    # 1-D, CE = 2
])
def test_rk4_backward_matches_autograd_through_the_oracle(mesh_dims, B, burgers, over):
    """Hand-written backward of the fused RK4 step (north_star items 3-4; csrc/ell_kernels.cuh: k_ell_bwd_rk4):
    parameter gradients against autograd through the oracle's RK4 (fp64 bar as everywhere else), input gradients
    against autograd as well.  The reference integrates with Euler only; RK4 is pinned by the oracle alone."""
    over = dict(over, ode_method="rk4")
    model, out, ref_out, data = _compare_with_oracle(mesh_dims, B, over=over, burgers=burgers, backward=True, seed=3)
    assert model.last_graph.tile_ptr is not None
    # input gradients (x_comp, uu) through the same kernel
    opt = synth.burgers_opt(mesh_dims, **over) if burgers else synth.default_opt(mesh_dims, **over)
    ds = synth.SyntheticDataset(len(mesh_dims), mesh_dims)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    m2 = cuda_model(ds, opt, ref.state_dict())
    d_cpu, d_gpu = data.clone(), data.clone().to("cuda")
    for d in (d_cpu, d_gpu):
        d.x_comp = d.x_comp.clone().requires_grad_(True)
        d.uu_tensor = d.uu_tensor.clone().requires_grad_(True)
    w = torch.randn(ref_out.shape, generator=torch.Generator().manual_seed(5))
    (ref(d_cpu) * w).sum().backward()
    (m2(d_gpu) * w.cuda()).sum().backward()
    assert util.rel_err(d_gpu.x_comp.grad, d_cpu.x_comp.grad) <= 2e-5
    assert util.rel_err(d_gpu.uu_tensor.grad, d_cpu.uu_tensor.grad) <= 2e-5


@pytest.mark.parametrize("mesh_dims,B,burgers,over,extra", [
    ((50, 50), 2, False, {"num_layers": 3}, {}),                  # 2500-node tiles: nine rows per node do not fit a CTA
    ((64, 64), 1, False, {"num_layers": 2}, {}),                  # beyond one CTA: cluster forward, streaming backward
    ((64, 64), 1, False, {"num_layers": 2, "share_conv": False}, {"gad_no_cluster": True}),
    ((12, 12), 3, False, {"num_layers": 4}, {"gad_force_stream": True}),
    ((3000,), 2, True, {"num_layers": 3}, {}),
])
def test_rk4_backward_on_the_streaming_kernels(mesh_dims, B, burgers, over, extra):
    """RK4 backward beyond the mesh-resident kernel (csrc/stream_ell.cu: wide_backward_rk4_t -- three stage recomputes
    and four vjp passes of the Euler backward kernels per step): parameter gradients against autograd through the
    oracle's RK4 on the fp64 bar, input gradients against autograd."""
    over = dict(over, ode_method="rk4")
    model, out, ref_out, data = _compare_with_oracle(mesh_dims, B, over=over, burgers=burgers, backward=True, seed=3,
                                                     **extra)
    assert model.last_graph.wide_in is not None                   # the streaming rows were built for the backward
    opt = synth.burgers_opt(mesh_dims, **over) if burgers else synth.default_opt(mesh_dims, **over)
    ds = synth.SyntheticDataset(len(mesh_dims), mesh_dims)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    m2 = cuda_model(ds, opt, ref.state_dict(), **extra)
    # input gradients against the oracle in fp64; the bar widens to 4 x the fp32 oracle's own deviation from fp64 where
    # that is larger (3000-node chains: signed sums, gad_testutil.fp64_grads_and_noise_floor has the argument)
    ref64 = oracle_model(ds, opt)
    ref64.load_state_dict(ref.state_dict())
    ref64 = ref64.double()
    d_cpu, d_gpu, d64 = data.clone(), data.clone().to("cuda"), data.clone()
    for k in d64.keys():
        v = getattr(d64, k)
        if torch.is_tensor(v) and v.dtype == torch.float32:
            setattr(d64, k, v.double())
    for d in (d_cpu, d_gpu, d64):
        d.x_comp = d.x_comp.clone().requires_grad_(True)
        d.uu_tensor = d.uu_tensor.clone().requires_grad_(True)
    w = torch.randn(ref_out.shape, generator=torch.Generator().manual_seed(5))
    (ref(d_cpu) * w).sum().backward()
    (ref64(d64) * w.double()).sum().backward()
    (m2(d_gpu) * w.cuda()).sum().backward()
    for name in ("x_comp", "uu_tensor"):
        g64 = getattr(d64, name).grad
        floor = util.rel_err(getattr(d_cpu, name).grad, g64)
        assert util.rel_err(getattr(d_gpu, name).grad, g64) <= max(2e-5, 4.0 * floor), (name, floor)


def test_rk4_streaming_backward_equals_the_mesh_resident_one():
    """Same batch, same weights: the streaming RK4 backward (gad_force_stream) against the one-launch mesh-resident
    kernel -- two independent implementations of the same adjoint."""
    md, B = (14, 14), 4
    over = {"ode_method": "rk4", "num_layers": 3}
    opt = synth.default_opt(md, **over)
    ds = synth.SyntheticDataset(2, md)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    data = synth.make_batch(md, B, seed=2)
    grads = []
    for extra in ({}, {"gad_force_stream": True}):
        m = cuda_model(ds, opt, ref.state_dict(), **extra)
        m.train()
        F.l1_loss(m(data), data.x_phys.cuda()).backward()
        grads.append(grads_of(m))
    scale = max(float(g.abs().max()) for g in grads[0].values())
    for k in grads[0]:
        assert float((grads[0][k] - grads[1][k]).abs().max()) <= 2e-5 * scale, k


def test_rk4_backward_refuses_what_it_cannot_do():
    opt = synth.default_opt((10, 10), ode_method="rk4", learn_step=True)
    m = cuda_model(synth.SyntheticDataset(2, (10, 10)), opt)
    out = m(synth.make_batch((10, 10), 2))
    with pytest.raises(NotImplementedError):
        out.sum().backward()


def test_config5_slice_50x50_batch64_fwd_bwd():
    model, *_ = _compare_with_oracle((50, 50), 64)
    assert model.last_graph.tile_ptr is not None and model.last_graph.max_tile_nodes == 2500


def test_config5_per_gpu_shard_through_replication():
    """cfg 5 at its per-GPU size on 8 GPUs (1024 meshes of 50x50 = 2.56 M nodes, 14.6 M edges), checked
    through a size-independent property: the batch is 128 copies of an 8-mesh batch, so every copy must come
    out bit for bit the same, equal the oracle on the 8 meshes, and (mean loss) give the 8-mesh gradient."""
    mesh_dims, B0, R = (50, 50), 8, 128
    opt = synth.default_opt(mesh_dims)
    ds = synth.SyntheticDataset(2, mesh_dims)
    small = synth.make_batch(mesh_dims, B0, seed=21)
    big = synth.make_batch(mesh_dims, B0 * R, seed=21)
    for k in ("x_comp", "x_phys", "f_tensor", "uu_tensor", "u_true_tensor"):
        v = getattr(small, k)
        setattr(big, k, v.repeat((R,) + (1,) * (v.dim() - 1)).contiguous())
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    ref_out = ref(small)
    gnn_oracle.mesh_loss(ref_out, small.x_phys).backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    model = cuda_model(ds, opt, ref.state_dict())
    model.train()
    out = model(big)
    assert model.last_graph.T == B0 * R and model.last_graph.max_tile_nodes == 2500
    o = out.detach().reshape(R, ref_out.shape[0], -1)
    assert torch.equal(o, o[0:1].expand_as(o))
    assert util.rel_err(o[0], ref_out) <= COORD_TOL
    F.l1_loss(out, big.x_phys.cuda()).backward()
    g64, floor, scale = util.fp64_grads_and_noise_floor(ds, opt, small, ref, ref_grads)
    util.check_grads_conditioned(grads_of(model), g64, floor, scale, tol=GRAD_TOL)


def test_scaled_weights_stress_softmax():
    # x4 weights: logits 16x larger -> near one-hot attention; parity is still within the bar
    _compare_with_oracle((24, 24), 8, wscale=4.0)


def test_variable_size_batch_and_tile_packing():
    sizes = [[5, 5], [9, 9], [30, 30], [7, 7], [12, 12], [6, 6], [20, 20]] * 3
    opt = synth.default_opt(sizes[0])
    ds = synth.SyntheticDataset(2, sizes[0])
    data = synth.make_mixed_batch(sizes, seed=3)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    ref_out = ref(data)
    gnn_oracle.mesh_loss(ref_out, data.x_phys).backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    outs = []
    for tile_nodes in (64, 512, 1024, 3000):
        model = cuda_model(ds, opt, ref.state_dict(), gad_tile_nodes=tile_nodes)
        model.train()
        out = model(data)
        assert util.rel_err(out, ref_out) <= COORD_TOL
        F.l1_loss(out, data.x_phys.cuda()).backward()
        check_grads(grads_of(model), ref_grads)
        outs.append(out.detach().cpu())
        if model.last_graph.tile_ptr is not None:
            tp = model.last_graph.tile_ptr.cpu().numpy()
            bounds = np.concatenate([[0], np.cumsum([a * b for a, b in sizes])])
            assert set(tp.tolist()) <= set(bounds.tolist())          # tiles never split a mesh
    for o in outs[1:]:
        assert util.rel_err(o, outs[0]) <= 1e-6      # the tiling does not change the arithmetic


def test_run_to_run_determinism():
    mesh_dims, B = (30, 30), 64
    opt = synth.default_opt(mesh_dims)
    ds = synth.SyntheticDataset(2, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=9)
    res = []
    for _ in range(3):
        torch.manual_seed(1)
        model = cuda_model(ds, opt)
        model.train()
        out = model(data)
        F.l1_loss(out, data.x_phys.cuda()).backward()
        res.append((out.detach().cpu(), grads_of(model)))
    for out, gr in res[1:]:
        assert torch.equal(out, res[0][0])
        for n in gr:
            assert torch.equal(gr[n], res[0][1][n]), n       # no floating-point atomics anywhere


def test_input_gradients_match_autograd():
    mesh_dims, B = (9, 9), 3
    opt = synth.default_opt(mesh_dims)
    ds = synth.SyntheticDataset(2, mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=2)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    rd = data.clone()
    for k in ("x_comp", "f_tensor", "uu_tensor"):
        getattr(rd, k).requires_grad_(True)
    gnn_oracle.mesh_loss(ref(rd), data.x_phys).backward()
    model = cuda_model(ds, opt, ref.state_dict())
    gd = data.clone().to("cuda")
    for k in ("x_comp", "f_tensor", "uu_tensor"):
        getattr(gd, k).requires_grad_(True)
    F.l1_loss(model(gd), gd.x_phys).backward()
    for k in ("x_comp", "f_tensor", "uu_tensor"):
        a, b = getattr(gd, k).grad.cpu(), getattr(rd, k).grad
        assert (a - b).abs().max().item() <= GRAD_TOL * max(b.abs().max().item(), 1e-12), k


# --------------------------------------------------------------------------------------
# operator seam and the loss helper
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("conv_type", ["GRAND_plus", "GRAND"])
@pytest.mark.parametrize("C", [8, 4, 2])
def test_operator_seam_single_layer(conv_type, C):
    md = (12, 12)
    opt = synth.default_opt(md, conv_type=conv_type, hidden_dim=C)
    data = synth.make_batch(md, 3, seed=5)
    ei = gnn_oracle.filtered_edge_index(data, opt, 2)
    N = data.x_comp.shape[0]
    torch.manual_seed(7)
    x = torch.randn(N, C)
    ref_cls = gnn_oracle.GRANDPlusConvRef if conv_type == "GRAND_plus" else gnn_oracle.GRANDConvRef
    ref = ref_cls(copy.deepcopy(opt), C, C, heads=1)
    from g_adaptivity_b200.GNN import get_conv
    conv = get_conv(copy.deepcopy(opt), conv_type, C, C, 8).cuda()
    conv.load_state_dict(ref.state_dict())
    xr = x.clone().requires_grad_(True)
    xg = x.clone().cuda().requires_grad_(True)
    args = (None, None) if conv_type == "GRAND_plus" else ()
    r_ref = ref(xr, ei, *args)
    r_gpu = conv(xg, ei.cuda(), *args)
    assert util.rel_err(r_gpu, r_ref) <= COORD_TOL
    assert (conv.stored_alpha.cpu() - ref.stored_alpha).abs().max().item() <= 2e-5
    assert torch.equal(conv.stored_ei.cpu(), ei)
    w = torch.randn(N, C)
    (r_ref * w).sum().backward()
    (r_gpu * w.cuda()).sum().backward()
    assert util.rel_err(xg.grad, xr.grad) <= GRAD_TOL
    got = {n: p.grad.cpu() for n, p in conv.named_parameters() if p.grad is not None}
    check_grads(got, {n: p.grad for n, p in ref.named_parameters() if p.grad is not None})


@pytest.mark.parametrize("kind", ["l1", "mse"])
def test_mesh_loss_kernel(kind):
    torch.manual_seed(0)
    a = torch.randn(100_003, 2, device="cuda")
    b = torch.randn(100_003, 2, device="cuda")
    loss, g = GF.mesh_loss(a, b, kind)
    ar = a.clone().requires_grad_(True)
    ref = F.l1_loss(ar, b) if kind == "l1" else F.mse_loss(ar, b)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (g - ar.grad).abs().max().item() <= 1e-9


def test_no_cpu_fallback():
    opt = synth.default_opt((6, 6))      # device='cpu'
    m = GNN(synth.SyntheticDataset(2, (6, 6)), opt)
    with pytest.raises(RuntimeError):
        m(synth.make_batch((6, 6), 1))
    conv = GRAND_plusConv(opt, 8, 8, heads=1, concat=False, bias=False, root_weight=False)
    with pytest.raises(RuntimeError):
        conv(torch.zeros(4, 8), torch.zeros((2, 0), dtype=torch.long))
